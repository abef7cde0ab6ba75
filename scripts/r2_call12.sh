#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c12
U3=1024 U4=640 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.tv_launches.csv python scripts/tv_breakdown.py > $O.tv_breakdown.log 2>&1; echo "ncu rc=$?"
tail -n 4 $O.tv_breakdown.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2c12.tv_launches.csv')) if len(r)>5]
hdr=None
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(d['Metric Value'].replace(',',''))
    u=d['Metric Unit']
    if u=='us': v*=1e3
    elif u=='ms': v*=1e6
    agg[d['Kernel Name'][:80]][0]+=1; agg[d['Kernel Name'][:80]][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print(f"{v[1]/1e6:9.3f} ms {100*v[1]/tot:5.1f}% n={v[0]:4d} {k}")
PY
