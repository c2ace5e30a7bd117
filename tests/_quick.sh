#!/bin/bash
# quick GPU check used during development: parity + C4/C5 stage times with replay debug counters
export FSLRC_DEBUG=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_sharded_emulated.py tests/test_gpu_options.py -x -q 2>&1 | tail -15
for c in C4 C5; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --e2e-depth 1 > gpurun_out/quick_$c.log 2>&1
  grep fslrc gpurun_out/quick_$c.log | tail -2
  python -c "
import json;j=json.loads(open('gpurun_out/quick_$c.log').read().strip().splitlines()[-1]);print('$c', round(j['ms_per_step'],2),{k:round(v,2) for k,v in j['stage_ms'].items()}); print(' e2e', round(j['e2e']['ms_per_step'],2), 'tests', j['pair_tests'], 'roof', j['roofline']['frac'], j['roofline']['whole_pair_stage']['frac'])" || tail -20 gpurun_out/quick_$c.log
done
