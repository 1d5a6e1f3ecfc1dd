#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench2.log 2>&1; echo "exit=$?" >> gpurun_out/t_bench2.log; tail -n 3 gpurun_out/t_bench2.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/t_dp2.log 2>&1; echo "exit=$?" >> gpurun_out/t_dp2.log; tail -n 6 gpurun_out/t_dp2.log
