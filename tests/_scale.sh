#!/bin/bash
# the driver's scaling sequence: N = 1, 2, 4, 8 back to back on one box
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n$n.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.log 2>&1
  fi
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r02_scale_n$n.log") if l.startswith("{")][-1])
    print("N=$n", round(j["ms_per_step"],2), "ms/step", round(j["value"]/1e6,1), "M reads/s; e2e", round(j["e2e"]["ms_per_step"],2), "ms", j.get("sharded_result_check"))
    print("   ", {k:round(v,2) for k,v in j["stage_ms"].items()})
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/r02_scale_n$n.log").read()[-1500:])
PY
done
FSLRC_DEBUG_MG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 6 --warmup 3 2>&1 | grep "^\[mg\]" | sed -n 4,7p | cut -c1-220
