#!/bin/bash
# weak-scaling bench on one multi-GPU box: bash tools/gpu_scale.sh "8 4"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
port=29600
for n in ${1:-8 4}; do
  port=$((port + 1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench$n.log 2>&1
  echo "exit=$?" >> gpurun_out/t_bench$n.log
  grep '^{' gpurun_out/t_bench$n.log | cut -c1-260
done
