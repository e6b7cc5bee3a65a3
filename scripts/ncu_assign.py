"""Minimal driver for ncu captures of the search kernel at BASELINE config 2 (run under gpurun + ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vector_quantize_pytorch_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, D, K = 262144, 256, 1024
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x = torch.randn(N, D, device=dev).bfloat16()
c = torch.randn(K, D, device=dev)
cb = ops.prepare_codebook(c, False)
for _ in range(3):
    r = ops.search(x, cb, c, n_passes=passes)
torch.cuda.synchronize()
print("flagged", r.flag_count.item())
