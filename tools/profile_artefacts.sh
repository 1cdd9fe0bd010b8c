#!/bin/bash
# Turn the captures of tools/gpu_profile.sh <tag> (gpurun_out/) into the tracked summaries under profiles/.
# usage: tools/profile_artefacts.sh <tag>     (run in the build container; the .o files must match the captured build)
TAG=${1:?tag}
python tools/ncu_summary.py $TAG
for k in spectrum:rg_spectrum.o:spectrum_kernelILi4Ed hist:rg_hist.o:hist_kernel nthcomp:rg_setup.o:nthcomp_kernel interp:rg_setup.o:nthcomp_interp_kernelILb0; do
  IFS=: read n o f <<< "$k"
  [ -f gpurun_out/prof_${n}_$TAG.ncu-rep ] || continue
  python tools/ncu_by_line.py gpurun_out/prof_${n}_$TAG.ncu-rep relagn_b200/_build/$o $f 60 > profiles/${TAG}_${n}_by_line.txt
  python tools/ncu_opcodes.py gpurun_out/prof_${n}_$TAG.ncu-rep 20 > profiles/${TAG}_${n}_opcodes.txt
done
for f in launches_$TAG.csv:${TAG}_launches.csv bench_$TAG.json:${TAG}_bench.json bench_reference_$TAG.json:${TAG}_bench_reference.json configs_$TAG.json:${TAG}_configs.json bench_8gpu_$TAG.json:${TAG}_bench_8gpu.json; do
  IFS=: read a b <<< "$f"; [ -f gpurun_out/$a ] && cp gpurun_out/$a profiles/$b
done
ls profiles | grep "^$TAG" | tr '\n' ' '
