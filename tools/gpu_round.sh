#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run t_all     900 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread
TAILN=3 run t_bench   900 python bench.py --steps 10 --warmup 3
run t_ncu_bwd 900 ncu --set full --clock-control none --import-source on -k regex:agg_bwd -s 20 -c 1 -o gpurun_out/prof_aggbwd_r01 python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline
run t_ncu_fwd 900 ncu --set full --clock-control none --import-source on -k regex:agg_fwd -s 20 -c 1 -o gpurun_out/prof_aggfwd_r01 python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline
