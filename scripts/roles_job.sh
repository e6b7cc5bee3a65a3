#!/bin/bash
# rebuild with the in-kernel cycle counters, print the per-role accounting, restore nothing (the box is scratch)
VQB_PROFILE=1 python -c "
from vector_quantize_pytorch_b200 import build
build.build(force=True)" > /dev/null 2>&1
timeout 120 python scripts/gpu_roles.py "$@"
