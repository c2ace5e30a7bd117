#!/bin/bash
# evidence for profiles/ (round 2, final code): launch list + ncu --set full of every kernel of ONE resident C4 step (the
# second one); the big report stays on the box, its raw page travels as CSV; a small report with sources for the hot kernels;
# the WALK replay of C5 (job board).  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
T=r02b
CMD="python bench.py --config C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1"
timeout 600 $CMD > gpurun_out/${T}_plain.log 2>&1 || { tail -5 gpurun_out/${T}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c4.csv $CMD > gpurun_out/${T}_ncu_list.log 2>&1
N=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${T}_launches_c4.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
names=[r[rows[hdr].index('Kernel Name')] for r in rows[hdr+1:]]
print(next(i for i,n in enumerate(names) if n.startswith('k_number'))+1)
PY
)
echo "launches per step: $N"
timeout 2400 ncu --set full --clock-control none -s $N -c $N -o /tmp/${T}_full_c4 -f $CMD > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i /tmp/${T}_full_c4.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_all_kernels_c4.csv 2>/dev/null
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_eval|k_hits|k_plist|k_union_entries|k_replay_list" -s 5 -c 5 -o gpurun_out/${T}_hot_c4 -f $CMD > gpurun_out/${T}_ncu_hot.log 2>&1
CMD5="python bench.py --config C5 --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1"
timeout 600 $CMD5 > gpurun_out/${T}_plain_c5.log 2>&1 || { tail -5 gpurun_out/${T}_plain_c5.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_replay" -s 1 -c 1 -o gpurun_out/${T}_walk_c5 -f $CMD5 > gpurun_out/${T}_ncu_walk_c5.log 2>&1
ls -la gpurun_out/ | grep ${T}
