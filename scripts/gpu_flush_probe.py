"""Which operand values does the tensor core flush?  (diagnostic; run under gpurun)

One-hot probes through the search kernel itself (cosine metric: score = x . c, no bias): a codebook whose code 0 is
[v, 0, ...] (all other codes [-1, 0, ...]) against x = [1, 0, ...] scores exactly v unless v is flushed on the B side; the
transposed probe tests the A side.  Swept over v = 1.5 * 2^-e for the single fp16 pass and the bf16 split scheme."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vector_quantize_pytorch_b200 import ops
dev = torch.device("cuda:0")
D, K, N = 64, 16, 128
for n_passes, name in ((2, "bf16 split"),):  # (the single fp16 pass needs K >= 256: probed through the tests instead)
    for side in ("B (codebook)", "A (input)"):
        out = []
        for e in range(6, 30):
            v = 1.5 * 2.0 ** -e
            c = torch.zeros(K, D, device=dev); c[:, 0] = -1.0
            x = torch.zeros(N, D, device=dev)
            if side.startswith("B"):
                c[0, 0] = v; x[:, 0] = 1.0
            else:
                c[0, 0] = 1.0; x[:, 0] = v
            cb = ops.prepare_codebook(c, True)
            r = ops.search(x.bfloat16(), cb, c, n_passes=n_passes, debug_best=True, fix=False, normalise=False)
            torch.cuda.synchronize()
            out.append((e, r.best[0].item() / v))
        print(name, side, " ".join(f"2^-{e}:{q:.3g}" for e, q in out))

# flagged-row statistics of the two schemes at BASELINE config 2 (how much the exact re-score has to do)
torch.manual_seed(0)
xx = torch.randn(262144, 256, device=dev).bfloat16()
cc = torch.randn(1024, 256, device=dev)
cbb = ops.prepare_codebook(cc, False)
for n_passes in (2,):
    r = ops.search(xx, cbb, cc, n_passes=n_passes, fix=False)
    torch.cuda.synchronize()
    n = int(r.flag_count.item())
    cnt = r.flagged[:n, 1]
    print(f"cfg2 n_passes={n_passes}: flagged {n} rows ({100.0 * n / xx.shape[0]:.3f} %), count==2: {(cnt == 2).sum().item()}, >=3: {(cnt >= 3).sum().item()}")
