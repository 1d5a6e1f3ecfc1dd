mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x --timeout 120 > gpurun_out/t_k.log 2>&1; tail -3 gpurun_out/t_k.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_fwd_pipe -s 3 -c 1 -f -o gpurun_out/prof_fwd python tools/bench_agg.py --which fwd --iters 5 > gpurun_out/ncu_fwd.log 2>&1
python tools/ncu_hot.py gpurun_out/prof_fwd.ncu-rep 40 > gpurun_out/fwd_hot.txt 2>&1
ncu -i gpurun_out/prof_fwd.ncu-rep --page raw --csv > gpurun_out/fwd_raw.csv 2>/dev/null
ls -la gpurun_out/prof_fwd.ncu-rep
