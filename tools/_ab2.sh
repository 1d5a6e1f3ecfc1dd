mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "tensor_core_large or agg_bwd" --timeout 60 > gpurun_out/t_ring.log 2>&1; tail -3 gpurun_out/t_ring.log
echo "default"; timeout 120 python tools/bench_agg.py --which bwd_fused,fwd --iters 30 2>&1 | tail -1
for v in ${VARIANTS:-}; do
echo "variant $v"; PB200_LIB=$PWD/gpurun_ab/lib_$v.so timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "tensor_core_large" --timeout 60 2>&1 | tail -1
PB200_LIB=$PWD/gpurun_ab/lib_$v.so timeout 120 python tools/bench_agg.py --which bwd_fused --iters 30 2>&1 | tail -1
done
