mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_scale_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/t_m.log 2>&1; tail -5 gpurun_out/t_m.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/b_cur.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/b_cur.log'):
    if l.startswith('{'):
        j=json.loads(l); print('cur', j['ms_per_step'], j['e2e']['ms_per_step'], j['clocks']['sm_mhz'], 'ours', j['our_kernels_ms_per_step'])
PY
