#!/bin/bash
# evidence for profiles/: launch list + ncu --set full of every kernel of ONE resident C4 step (the second one);
# the big report stays on the box, its raw page travels as CSV; a small report with sources for the hot kernels
mkdir -p gpurun_out
CMD="python bench.py --config C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1"
timeout 600 $CMD > gpurun_out/r02_plain.log 2>&1 || { tail -5 gpurun_out/r02_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
N=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c4.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
names=[r[rows[hdr].index('Kernel Name')] for r in rows[hdr+1:]]
print(next(i for i,n in enumerate(names) if n.startswith('k_number'))+1)
PY
)
echo "launches per step: $N"
timeout 2400 ncu --set full --clock-control none -s $N -c $N -o /tmp/r02_full_c4 -f $CMD > gpurun_out/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_full_c4.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_all_kernels_c4.csv 2>/dev/null
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_eval|k_hits|k_plist|k_union_entries|k_replay" -s 5 -c 5 -o gpurun_out/r02_hot_c4 -f $CMD > gpurun_out/r02_ncu_hot.log 2>&1
ls -la gpurun_out/
