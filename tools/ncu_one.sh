#!/bin/bash
# ncu --set full capture of one kernel family inside one bench step:  tools/ncu_one.sh <name> <kernel regex> [skip] [count]
name=$1; rx=$2; skip=${3:-8}; cnt=${4:-1}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -s $skip -c $cnt -f \
  -o gpurun_out/prof_${name}_r02 python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline --profiler-range > gpurun_out/ncu_$name.log 2>&1
ls -la gpurun_out/prof_${name}_r02.ncu-rep
