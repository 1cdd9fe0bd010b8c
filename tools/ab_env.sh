#!/bin/bash
# A/B on the GPU box with environment knobs: usage: tools/ab_env.sh "VAR=1" "VAR=2 OTHER=x" ...
for F in "$@"; do
  env $F python bench.py --steps 3 --warmup 2 --no-cpu --no-fit 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print('ENV [$F] value=%.0f ms=%.1f e2e=%.0f spectrum=%.1f hist=%.1f nthcomp=%.1f setup=%.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], k['spectrum_kernel_ms'], k['hist_kernel_ms'], k['nthcomp_kernel_ms'], k['setup_kernel_ms']))"
done
