mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 120 > gpurun_out/t_k.log 2>&1; tail -3 gpurun_out/t_k.log
for r in 0 1; do echo "PIPE=$r"; PB200_AGG_FWD_PIPE=$r timeout 120 python tools/bench_agg.py --which fwd --iters 30 2>&1 | tail -1; done
