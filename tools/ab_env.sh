#!/bin/bash
# tools/ab_env.sh VAR v1 v2 ...: bench.py (no secondary, no CPU arm) once per value of an environment switch
var=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $var=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/ab_${var}_$v.log 2>&1
  python - <<PY
import json
for l in open('gpurun_out/ab_${var}_$v.log'):
    if l.startswith('{'):
        j = json.loads(l); print('$var=$v', 'value ms', round(j['ms_per_step'], 2), 'e2e ms', round(j['e2e']['ms_per_step'], 2), 'sm', j['clocks']['sm_mhz'])
PY
done
