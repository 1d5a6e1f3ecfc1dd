#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run t_all     900 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread
TAILN=2 run t_smoke   300 python __graft_entry__.py --smoke
TAILN=3 run t_bench   900 python bench.py --steps 10 --warmup 3
TAILN=45 run t_prof    300 python tools/profile_step.py bf16
