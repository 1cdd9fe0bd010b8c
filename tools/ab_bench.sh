#!/bin/bash
# A/B on the GPU box: rebuild the library with different -D flags and run a short bench for each.
# usage: tools/ab_bench.sh "<flags A>" "<flags B>" ...
for F in "$@"; do
  RG_NVCC_EXTRA="$F" python -m relagn_b200.build --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  python bench.py --steps 2 --warmup 2 --no-cpu --sets 65536 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('FLAGS [$F] value=%.0f e2e=%.0f spectrum_ms=%.1f nthcomp_ms=%.1f setup_ms=%.1f fp64frac=%.3f' % (d['value'], d['e2e']['value'], d['kernels']['spectrum_kernel_ms'], d['kernels']['nthcomp_kernel_ms'], d['kernels']['setup_kernel_ms'], d['fp64_roofline']['frac']))"
done
RG_NVCC_EXTRA="" python -m relagn_b200.build --force > /dev/null 2>&1
