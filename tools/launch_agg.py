import csv,collections,re,sys
lines=open(sys.argv[1]).read().splitlines()
i0=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.DictReader(lines[i0:]))
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    n=re.sub(r"\(.*","",r["Kernel Name"])[:90]+" grid="+r["Grid Size"]
    agg[n][0]+=1; agg[n][1]+=float(r["Metric Value"])/1e3
tot=sum(v[1] for v in agg.values())
print(len(rows),"launches",tot,"us")
for n,v in sorted(agg.items(),key=lambda x:-x[1][1])[:28]: print(f"{v[1]:9.1f} us {v[0]:4d} {v[1]/v[0]:8.1f} avg  {n}")
