#!/bin/bash
# ncu evidence for profiles/: launch list of ONE timed step + one --set full capture per dominant kernel.
# Numbers printed by bench.py under ncu are never bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline --no-secondary --profiler-range"
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; }
run t_launch ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $B
for spec in ${SPECS:-aggfwd:agg_fwd_pipe_kernel aggbwdring:agg_bwd_ring_kernel gemm:gemm_tcgen05_kernel bn:bn_ dropbits:dropout_bits ce:ce_rows_kernel chord:chord_embed pool:bar_ rows:rows_permute}; do
  name=${spec%%:*}; rx=${spec##*:}; cnt=1; skip=8
  case $name in ce|chord|pool|rows) skip=0;; esac
  [ $name = rows ] && cnt=2
  [ $name = gemm ] && cnt=3
  [ $name = bn ] && cnt=5
  [ $name = ce ] && cnt=2
  [ $name = chord ] && cnt=2
  [ $name = pool ] && cnt=4
  run t_ncu_$name ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/prof_${name}_$TAG $B
done
# summaries are made here (the reports themselves exceed what gpurun copies back); keep the backward's report for source-level work
python tools/summarize_profiles.py $TAG gpurun_out/summaries_$TAG > gpurun_out/summarize.log 2>&1
ls -la gpurun_out/*.ncu-rep
find gpurun_out -name '*.ncu-rep' ! -name 'prof_aggbwdring_*' -delete
du -sh gpurun_out
