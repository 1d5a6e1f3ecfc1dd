mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/b_cur.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/b_cur.log'):
    if l.startswith('{'):
        j=json.loads(l); print('cur', j['ms_per_step'], j['e2e']['ms_per_step'], j['clocks']['sm_mhz'], 'ours', j['our_kernels_ms_per_step'])
        for k in j['kernels']:
            if 'agg' in k['kernel']: print('   ', k['kernel'], round(k['avg_ms'],3), k.get('frac') and round(k['frac'],3))
PY
done
