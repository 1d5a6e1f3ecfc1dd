#!/bin/bash
# One gpurun call: run the GPU test tiers in separate processes (a hang in one must not hide the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run t_diag    200 python tools/diag_gemm.py
run t_graph   300 python -m pytest tests/test_graph_gpu.py -q -m gpu --timeout 120 --timeout-method=thread
run t_kern    400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not gemm" --timeout 120 --timeout-method=thread
run t_gemm    400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" --timeout 120 --timeout-method=thread
run t_model   600 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 200 --timeout-method=thread
run t_smoke   300 python __graft_entry__.py --smoke
run t_bench   900 python bench.py --steps 5 --warmup 3
