#!/bin/bash
# one point of the driver's scaling sequence (N ranks over NCCL on one box), plus the phase times of the exchange steps
mkdir -p gpurun_out
n=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02b_scale_n$n.log 2>&1
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r02b_scale_n$n.log") if l.startswith("{")][-1])
    print("N=$n", round(j["ms_per_step"],2), "ms/step", round(j["value"]/1e6,1), "M reads/s; e2e", round(j["e2e"]["ms_per_step"],2), "ms", j.get("sharded_result_check"))
    print("   ", {k:round(v,2) for k,v in j["stage_ms"].items()})
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/r02b_scale_n$n.log").read()[-1500:])
PY
FSLRC_DEBUG_MG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29650+n)) bench.py --gpus $n --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep "^\[mg\]" | sed -n 8,10p | cut -c1-220 | tee gpurun_out/r02b_mg_phases_n$n.log
