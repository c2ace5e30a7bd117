#!/bin/bash
# quick GPU check used during development: parity + C4/C5 stage times with replay debug counters
export FSLRC_DEBUG=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for c in C4 C5; do
  timeout 300 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/quick_$c.log 2>&1
  grep fslrc gpurun_out/quick_$c.log | tail -2
  python -c "
import json;j=json.loads(open('gpurun_out/quick_$c.log').read().strip().splitlines()[-1]);print('$c', round(j['ms_per_step'],2),{k:round(v,2) for k,v in j['stage_ms'].items()})"
done
