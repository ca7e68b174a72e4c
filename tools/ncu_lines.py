#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.

usage: ncu -i rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_lines.py src.csv [warp_windows]
"""
import csv, sys, collections

def main():
    path = sys.argv[1]
    ww = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(path, newline="")))
    cur_file = None
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or cur_file is None:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr, r))
        # the first "Source" column is the CUDA line, later one is SASS; dict keeps the last -> use indexes
        src = r[1]
        inst = d.get("Instructions Executed", "") or "0"
        smp = d.get("# Samples", "") or "0"
        stall = d.get("Warp Stall Sampling (All Samples)", "") or "0"
        key = (cur_file, ln)
        e = per.setdefault(key, [src, 0, 0, 0])
        # SASS rows have an empty line number and were skipped above: these are the CUDA-line aggregates
        try:
            e[1] += int(float(inst)); e[2] += int(float(smp))
        except ValueError:
            pass
    tot = sum(e[1] for e in per.values())
    tots = sum(e[2] for e in per.values())
    print("total inst", tot, "samples", tots)
    items = sorted(per.items(), key=lambda kv: -kv[1][1])
    for (f, ln), e in items[:45]:
        extra = f"{e[1] / ww:8.1f}/ww" if ww else ""
        print(f"{f[:16]:16s}:{ln:4d} {e[1] / tot * 100:5.1f}% inst {extra} smp {e[2] / max(tots,1) * 100:5.1f}%  {e[0].strip()[:110]}")

if __name__ == "__main__":
    main()
