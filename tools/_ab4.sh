mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_scale_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/t_m.log 2>&1; tail -5 gpurun_out/t_m.log
for m in 0 1; do
PB200_SPLIT_HEADS=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/b_split$m.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/b_split$m.log'):
    if l.startswith('{'):
        j=json.loads(l); print('split=$m', j['ms_per_step'], j['e2e']['ms_per_step'], j['clocks']['sm_mhz'])
PY
done
