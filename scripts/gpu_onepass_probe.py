"""How many rows would a hi-only single pass hand to the exact re-score?  (diagnostic; run under gpurun)
BASELINE config 2 shape, randn inputs; codebook = randn (cold) and after 30 EMA steps (warm)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vector_quantize_pytorch_b200 as vqb
from vector_quantize_pytorch_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
xx = torch.randn(262144, 256, device=dev).bfloat16()
vq = vqb.VectorQuantize(dim=256, codebook_size=1024).to(dev)
vq.train()
def probe(tag, cc):
    cbb = ops.prepare_codebook(cc, False)
    for n_passes in (2, 1):
        r = ops.search(xx, cbb, cc, n_passes=n_passes, fix=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            r = ops.search(xx, cbb, cc, n_passes=n_passes, fix=False)
        e1.record(); torch.cuda.synchronize()
        n = int(r.flag_count.item()); nb = int(r.rescan_count.item())
        cnt = r.flagged[:n, 1]
        print(f"{tag} n_passes={n_passes}: search {e0.elapsed_time(e1) / 10 * 1000:.1f} us, 2-3 candidates {n} rows ({100.0 * n / xx.shape[0]:.3f} %) "
              f"[2: {(cnt == 2).sum().item()}, 3: {(cnt == 3).sum().item()}], >3 (rescan) {nb} rows ({100.0 * nb / xx.shape[0]:.3f} %)")
probe("cold", vq._codebook.embed[0].clone())
for _ in range(30):
    vq(xx.view(64, 4096, 256))
probe("warm30", vq._codebook.embed[0].clone())
x2 = (torch.randn(262144, 256, device=dev) * torch.rand(262144, 1, device=dev) * 3).bfloat16()   # mixed norms
xx = x2
probe("warm30/mixed-norm x", vq._codebook.embed[0].clone())
