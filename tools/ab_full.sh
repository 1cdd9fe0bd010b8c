#!/bin/bash
# A/B on the GPU box at the bench's full size (125 000 sets): each argument is "ENV=.. ENV=.." and/or "-DFLAG=.." words;
# -D words rebuild the library first.  usage: tools/ab_full.sh "X=1" "-DRG_NTH_PREFETCH=0" "RG_SLAB_MIN=16384"
LAST=""
for F in "$@"; do
  DEFS=$(for w in $F; do case $w in -D*) echo -n "$w ";; esac; done)
  ENVS=$(for w in $F; do case $w in -D*) ;; *) echo -n "$w ";; esac; done)
  if [ "$DEFS" != "$LAST" ]; then
    RG_NVCC_EXTRA="$DEFS" python -m relagn_b200.build --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
    LAST="$DEFS"
  fi
  env $ENVS python bench.py --steps 3 --warmup 3 --no-cpu --no-fit 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print('AB [$F] value=%.0f e2e=%.0f ms=%.1f | spectrum %.1f nthcomp %.1f setup %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], k['spectrum_kernel_ms'], k['nthcomp_kernel_ms'], k['setup_kernel_ms']), {x:round(v,1) for x,v in k.items() if isinstance(v,(int,float))})"
done
if [ -n "$LAST" ]; then RG_NVCC_EXTRA="" python -m relagn_b200.build --force > /dev/null 2>&1; fi
