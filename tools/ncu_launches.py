import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki][:70]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]: print(f"{n:4d} {t/1e6:9.3f} ms total {t/n/1e3:9.1f} us avg  {k}")
