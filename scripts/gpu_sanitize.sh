#!/bin/bash
# memcheck a small search-only launch and one golden replay (diagnostic; run under gpurun)
cat > /tmp/san_small.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vector_quantize_pytorch_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (N, D, K, dt) in ((128, 32, 64, torch.float32), (4096, 256, 1024, torch.bfloat16)):
    x = torch.randn(N, D, device=dev).to(dt)
    c = torch.randn(K, D, device=dev)
    cb = ops.prepare_codebook(c, False)
    out = ops.search(x, cb, c, fix=False)
    torch.cuda.synchronize()
    print("ok", N, D, K, dt)
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san_small.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in \|^=========                in " | head -60
