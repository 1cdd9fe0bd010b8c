#!/bin/bash
# A/B on the GPU box with compile-time knobs: rebuild the library with each flag set and run a short bench.
# usage: tools/ab_build.sh "<nvcc flags A>" "<nvcc flags B>" ...
for F in "$@"; do
  RG_NVCC_EXTRA="$F" python -m relagn_b200.build --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  python bench.py --steps 3 --warmup 2 --no-cpu --no-fit 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print('FLAGS [$F] value=%.0f ms=%.1f e2e=%.0f spectrum=%.1f hist=%.1f nthcomp=%.1f setup=%.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], k['spectrum_kernel_ms'], k['hist_kernel_ms'], k['nthcomp_kernel_ms'], k['setup_kernel_ms']))"
done
RG_NVCC_EXTRA="" python -m relagn_b200.build --force > /dev/null 2>&1
