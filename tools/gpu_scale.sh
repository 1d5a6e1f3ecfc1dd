#!/bin/bash
# tools/gpu_scale.sh N: the driver's bench invocation at N GPUs (reference arm first at N=1), logs under gpurun_out/
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_1gpu.log 2>&1
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_1gpu.log 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2>&1
fi
echo "exit=$?"; tail -c 400 gpurun_out/bench_${N}gpu.log
