#!/bin/bash
# development: the same C4 step with several tuning builds of the library (FSLR_B200_LIB)
mkdir -p gpurun_out
for v in "" a b c d e f g h; do
  lib=fslr_b200/csrc/libfslr_b200${v:+_$v}.so
  [ -f $lib ] || continue
  for c in ${CONFIGS:-C4}; do
    FSLR_B200_LIB=$PWD/$lib timeout 300 python bench.py --config $c --steps 5 --warmup 2 --no-cpu-baseline --e2e-depth 1 > gpurun_out/var_${v:-base}_$c.log 2>&1
    python -c "
import json;j=json.loads(open('gpurun_out/var_${v:-base}_$c.log').read().strip().splitlines()[-1]);print('${v:-base}', '$c', round(j['ms_per_step'],2),{k:round(v,2) for k,v in j['stage_ms'].items()})" || tail -3 gpurun_out/var_${v:-base}_$c.log
  done
done
