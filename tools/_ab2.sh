mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "tensor_core_large or agg_bwd" --timeout 60 > gpurun_out/t_ring.log 2>&1; tail -15 gpurun_out/t_ring.log
for r in 0 1; do
echo "RING=$r"; PB200_AGG_BWD_RING=$r timeout 120 python tools/bench_agg.py --which bwd_fused --iters 30 2>&1 | tail -3
done
