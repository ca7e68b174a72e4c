#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from an ncu source page export.

usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv
       python tools/ncu_lines.py src.csv [warp_windows] [--by-samples]
"""
import csv, sys


def num(x):
    try:
        return int(float(x))
    except (ValueError, TypeError):
        return 0


def main():
    path = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    ww = float(args[0]) if args else None
    by_samples = "--by-samples" in sys.argv
    per, cur, hdr = {}, None, None
    for r in csv.reader(open(path, newline="")):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or cur is None:
            continue
        try:
            ln = int(r[0])          # SASS rows have an empty line number: only the CUDA-line aggregates are used
        except ValueError:
            continue
        d = dict(zip(hdr, r))
        e = per.setdefault((cur, ln), [r[1], 0, 0])
        e[1] += num(d.get("Instructions Executed"))
        e[2] += num(d.get("# Samples"))
    tot = max(sum(e[1] for e in per.values()), 1)
    tots = max(sum(e[2] for e in per.values()), 1)
    print("total inst", tot, "samples", tots)
    key = (lambda kv: -kv[1][2]) if by_samples else (lambda kv: -kv[1][1])
    for (f, ln), e in sorted(per.items(), key=key)[:45]:
        extra = f"{e[1] / ww:8.1f}/ww" if ww else ""
        print(f"{f[:16]:16s}:{ln:4d} {e[1] / tot * 100:5.1f}% inst {extra} smp {e[2] / tots * 100:5.1f}%  {e[0].strip()[:110]}")


if __name__ == "__main__":
    main()
