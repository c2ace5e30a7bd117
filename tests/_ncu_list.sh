#!/bin/bash
# per-kernel launch list of one C4 step (cold-cache, serialised: compare shares, not absolutes)
mkdir -p gpurun_out
CFG=${1:-C4}
timeout 600 python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/plain_$CFG.log 2>&1 || { tail -5 gpurun_out/plain_$CFG.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$CFG.csv \
   python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/ncu_$CFG.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$CFG.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); idi=h.index('ID')
data=[(int(r[idi]), r[ki], float(r[vi].replace(',',''))) for r in rows[hdr+1:] if r[idi].isdigit()]
# keep the first resident step: from the first k_rows_fast/k_fill to the first k_number
names=[d[1] for d in data]
start=0
end=next(i for i,n in enumerate(names) if n.startswith('k_number'))
agg=collections.OrderedDict()
for _,n,v in data[start:end+1]:
    k=n.split('(')[0][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v/1e3
tot=sum(a[1] for a in agg.values())
print('first step: %d launches, %.1f us'%(end+1-start, tot))
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print('%8.1f us %5.1f%% x%-3d %s'%(t,100*t/tot,c,k))
PY
