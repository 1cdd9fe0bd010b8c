#!/bin/bash
# Run on the GPU box (gpurun): launch list + one full ncu capture of the dominant kernel, per B200_PROFILING.md.
# usage: tools/gpu_profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r1}; shift || true
CMD="python bench.py --sets 16384 --steps 1 --warmup 3 --no-cpu $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum_kernel -s 3 -c 1 -o gpurun_out/prof_spectrum_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nthcomp_kernel -s 3 -c 1 -o gpurun_out/prof_nthcomp_$TAG -f $CMD > gpurun_out/ncu_full2_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 3 -c 1 -o gpurun_out/prof_hist_$TAG -f $CMD > gpurun_out/ncu_full3_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nthcomp_interp -s 3 -c 1 -o gpurun_out/prof_interp_$TAG -f $CMD > gpurun_out/ncu_full4_$TAG.log 2>&1
ls -la gpurun_out/
