#!/bin/bash
# bisect on the GPU box with compile-time knobs: rebuild with each flag set and run one pytest selection
# usage: tools/ab_test.sh "<pytest -k expression>" "<nvcc flags A>" "<nvcc flags B>" ...
K="$1"; shift
for F in "$@"; do
  RG_NVCC_EXTRA="$F" python -m relagn_b200.build --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  echo "FLAGS [$F]: $(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -1)"
done
RG_NVCC_EXTRA="" python -m relagn_b200.build --force > /dev/null 2>&1
