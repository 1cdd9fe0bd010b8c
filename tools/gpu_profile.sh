#!/bin/bash
# Run on the GPU box (gpurun): launch list + one full ncu capture of each hot kernel, per B200_PROFILING.md.
# usage: [KERNELS="spectrum nthcomp hist interp"] tools/gpu_profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r1}; shift || true
KERNELS=${KERNELS:-spectrum nthcomp hist interp}
CMD="python bench.py --sets 16384 --steps 1 --warmup 3 --no-cpu $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
for k in $KERNELS; do
  case $k in
    spectrum) re=spectrum_kernel;; nthcomp) re=nthcomp_kernel;; hist) re=hist_kernel;; interp) re=nthcomp_interp;; *) re=$k;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:$re -s 3 -c 1 -o gpurun_out/prof_${k}_$TAG -f $CMD > gpurun_out/ncu_full_${k}_$TAG.log 2>&1
done
ls -la gpurun_out/
