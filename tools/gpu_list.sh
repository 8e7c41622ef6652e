#!/bin/bash
# ncu launch list (durations only) of a small bench run; prints per-kernel totals of the device-resident step
TAG=${TAG:-l}
PAIRS=${PAIRS:-256}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --pairs $PAIRS --no-cpu-baseline --no-ba > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
python - <<PY
import csv, collections, re
rows=list(csv.reader(l for l in open("gpurun_out/${TAG}_launches.csv") if l.startswith('"')))
h=rows[0]; ki,vi,ui,gi=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit'),h.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r'<.*','',r[ki].split('(')[0].replace('void ','').replace('adb::',''))
    v=float(r[vi].replace(',',''))*({'ns':1e-3,'us':1,'ms':1e3}.get(r[ui],1))
    k=(n,r[gi])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for (n,g),(c,t) in agg.items():
    if t/c>3: print(f"{n:32s} grid {g:>22s} x{c:3d} avg {t/c:9.1f} us")
PY
