#!/bin/bash
# secondary BASELINE.json configurations (tools/sweep.py) -> gpurun_out/sweep_*.jsonl
mkdir -p gpurun_out
for name in ${1:-layer batch generate}; do
  timeout 900 python tools/sweep.py $name > gpurun_out/sweep_$name.jsonl 2> gpurun_out/sweep_$name.err
  echo "$name exit=$?"; tail -n 3 gpurun_out/sweep_$name.err | cut -c1-300; wc -l gpurun_out/sweep_$name.jsonl
done
