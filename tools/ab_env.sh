#!/bin/bash
# A/B on the GPU box over environment settings: tools/ab_env.sh "VAR=1" "VAR=2" ...
for F in "$@"; do
  env $F python bench.py --steps 2 --warmup 2 --no-cpu --sets 65536 $BENCH_EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ENV [$F] value=%.0f e2e=%.0f spectrum_ms=%.1f nthcomp_ms=%.1f fp64frac=%.3f' % (d['value'], d['e2e']['value'], d['kernels']['spectrum_kernel_ms'], d['kernels']['nthcomp_kernel_ms'], d['fp64_roofline']['frac']))"
done
