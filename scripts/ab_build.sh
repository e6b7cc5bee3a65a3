#!/bin/bash
# A/B timing of build-flag variants of the search kernel (diagnostic helper): ab_build.sh "<flags A>" "<flags B>" ...
for flags in "$@"; do
  VQB_NVCC_EXTRA="$flags" python -c "
from vector_quantize_pytorch_b200 import build
build.build(force=True)" > /dev/null 2>&1
  for i in 1 2; do
    VQB_BENCH_SKIP_E2E=1 timeout 200 python bench.py --steps 50 --warmup 10 2>&1 | tail -1 | sed "s/^/[$flags] /"
  done
done
