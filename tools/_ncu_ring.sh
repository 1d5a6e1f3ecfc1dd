mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_bwd_ring -s 3 -c 1 -f -o gpurun_out/prof_ring python tools/bench_agg.py --which bwd_fused --iters 5 > gpurun_out/ncu_ring.log 2>&1
python tools/ncu_hot.py gpurun_out/prof_ring.ncu-rep 45 > gpurun_out/ring_hot.txt 2>&1
ncu -i gpurun_out/prof_ring.ncu-rep --page raw --csv > gpurun_out/ring_raw.csv 2>/dev/null
ls -la gpurun_out/prof_ring.ncu-rep
tail -50 gpurun_out/ring_hot.txt
