#!/bin/bash
# One GPU round: full -m gpu suite, smoke, bench. Logs under gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
TAILN=${TAILT:-25} run t_all     1500 python -m pytest tests -q -m gpu --timeout 600 --timeout-method=thread ${PYTEST_ARGS:-}
TAILN=2 run t_smoke   300 python __graft_entry__.py --smoke
TAILN=3 run t_bench   900 python bench.py --steps 10 --warmup 3
