mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/b2_$tag.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/b2_$tag.log'):
    if l.startswith('{'):
        j=json.loads(l); print('$tag', round(j['ms_per_step'],2), round(j['e2e']['ms_per_step'],2), j['clocks']['sm_mhz'])
PY
}
run mb8 PB200_BUCKET_MB=8
run mb32 PB200_BUCKET_MB=32
run mb128 PB200_BUCKET_MB=128
run ch4 NCCL_MAX_NCHANNELS=4
