import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_util import Golden
from vector_quantize_pytorch_b200 import ops
import vector_quantize_pytorch_b200 as vqb
from oracle import vq_oracle as O
dev = "cuda:0"
g = Golden("vq_readme_fp32")
x = torch.from_numpy(g["s0_x"]).to(dev)
e = torch.from_numpy(g["s0_pre_cb0_embed"]).to(dev).contiguous()
ref = g["s0_indices"].reshape(-1)
flat = x.reshape(-1, 256).contiguous()
cb = ops.prepare_codebook(e, False)
for trial in range(3):
    r = ops.search(flat, cb, e, debug_best=True)
    torch.cuda.synchronize()
    idx = r.idx.cpu().numpy()
    mm = np.nonzero(idx != ref)[0]
    print("plain search: mismatches", mm, "flagged", r.flag_count.item(), r.flagged[:r.flag_count.item()].cpu().numpy().tolist())
    q = torch.empty_like(flat); i64 = torch.empty(flat.shape[0], dtype=torch.int64, device=dev); ls = torch.zeros(1, dtype=torch.float64, device=dev)
    r2 = ops.search(flat, cb, e, fused=dict(q_out=q, idx64_out=i64, loss_sum=ls))
    torch.cuda.synchronize()
    mm2 = np.nonzero(i64.cpu().numpy() != ref)[0]
    print("fused search: mismatches(idx64)", mm2, " idx32 mism", np.nonzero(r2.idx.cpu().numpy() != ref)[0])
    for m in mm2:
        s = O.scores(flat[m:m+1].cpu().numpy(), e.cpu().numpy(), False)[0]
        order = np.argsort(-s)[:3]
        print("  row", m, "ours", i64[m].item(), "ref", ref[m], "top3", order, s[order])
vq = vqb.VectorQuantize(dim=256, codebook_size=512).to(dev)
with torch.no_grad():
    vq._codebook.embed.copy_(e[None]); vq._codebook.embed_avg.copy_(torch.from_numpy(g["s0_pre_cb0_embed_avg"]).to(dev)[None])
    vq._codebook.cluster_size.copy_(torch.from_numpy(g["s0_pre_cb0_cluster_size"]).to(dev)[None])
qq, ii, ll = vq(x)
torch.cuda.synchronize()
print("module: mismatches", np.nonzero(ii.cpu().numpy().reshape(-1) != ref)[0], "loss", ll.item(), "ref loss", g["s0_loss"])
