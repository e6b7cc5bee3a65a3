"""Per-role cycle accounting of vq_assign_kernel (diagnostic; run under gpurun)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vector_quantize_pytorch_b200 import ops, _C
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, D, K = 262144, 256, 1024
x = torch.randn(N, D, device=dev).bfloat16()
c = torch.randn(K, D, device=dev)
cb = ops.prepare_codebook(c, False)
MODES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(2, 0), (2, 1), (1, 0)]
for passes, mode in MODES:
    _C.lib.vqb_debug_set_mode(mode)
    for _ in range(3):
        ops.search(x, cb, c, n_passes=passes, fix=False)
    prof = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    _C.lib.vqb_debug_set_profile_buffer(prof.data_ptr())
    ops.search(x, cb, c, n_passes=passes, fix=False)
    torch.cuda.synchronize()
    _C.lib.vqb_debug_set_profile_buffer(None)
    p = prof.cpu().double()
    names = ["prod.wait_b_empty", "prod.total", "mma.wait_t_empty", "mma.wait_b_full", "mma.wait_x_full", "mma.wait_a_full",
             "mma.total", "epi0.wait_norms", "epi0.wait_t_full", "epi0.work", "epi0.merge", "epi0.total", "epi1.wait_t_full", "epi1.work",
             "epi1.merge", "epi1.total"]
    lead = p[0::2].mean(0); foll = p[1::2].mean(0)
    print(f"passes={passes} dbg_mode={mode} (kcycles, mean over CTAs; leader | follower)")
    for i, n in enumerate(names):
        print(f"  {n:18s} {lead[i]/1e3:9.1f} | {foll[i]/1e3:9.1f}")

_C.lib.vqb_debug_set_mode(0)
