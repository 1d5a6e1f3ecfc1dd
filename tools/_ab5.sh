mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_scale_gpu.py -q -m gpu -k two_gpu --timeout 280 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/b2_$tag.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/b2_$tag.log'):
    if l.startswith('{'):
        j=json.loads(l); print('$tag', round(j['ms_per_step'],2), round(j['e2e']['ms_per_step'],2), j['clocks']['sm_mhz'])
PY
}
run r1 A=1
run r2 A=1
run r3 A=1
