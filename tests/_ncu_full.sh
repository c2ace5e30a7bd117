#!/bin/bash
# ncu --set full of the pair-stage kernels, the union kernel and the replay (second resident step of a C4 run)
mkdir -p gpurun_out
CFG=${1:-C4}; PAT=${2:-"k_eval|k_hits|k_plist|k_union_entries|k_replay"}; N=${3:-5}
timeout 600 python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/plain_full_$CFG.log 2>&1 || { tail -5 gpurun_out/plain_full_$CFG.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $N -c $N -o gpurun_out/full_$CFG -f \
   python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/ncu_full_$CFG.log 2>&1
ls -la gpurun_out/full_$CFG.ncu-rep
