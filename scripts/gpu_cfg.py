"""A few forwards of one BASELINE.json config (for `ncu` launch lists): gpu_cfg.py {2|3|4|5}"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vector_quantize_pytorch_b200 as vqb
dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "4"
if which == "5":   # GroupedResidualVQ G=2 Q=8 (BASELINE configs[4], one rank's shard), fp32
    import time
    g = vqb.GroupedResidualVQ(dim=256, groups=2, num_quantizers=8, codebook_size=1024).to(dev)
    x = torch.randn(8, 4096, 256, device=dev)
    g.train()
    for _ in range(3):
        q, i, l = g(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        q, i, l = g(x)
    t1 = time.perf_counter()          # host time to ENQUEUE 10 forwards
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("ok grvq host_enqueue_ms %.3f total_ms %.3f" % ((t1 - t0) * 100, (t2 - t0) * 100))
    sys.exit(0)
if which == "3":   # ResidualVQ Q=8 shared codebook, bf16 (BASELINE configs[2] shape)
    rvq = vqb.ResidualVQ(dim=256, num_quantizers=8, codebook_size=1024, shared_codebook=True).to(dev)
    x = torch.randn(32, 8192, 256, device=dev).bfloat16()
    with torch.no_grad():
        e = torch.randn_like(rvq.layers[0]._codebook.embed)
        rvq.layers[0]._codebook.embed.copy_(e); rvq.layers[0]._codebook.embed_avg.copy_(e)
    rvq.train()
    for _ in range(3):
        q, i, l = rvq(x)
    torch.cuda.synchronize()
    print("ok rvq")
    sys.exit(0)
if which == "4":
    vq = vqb.VectorQuantize(dim=512, codebook_size=16384, use_cosine_sim=True).to(dev)
    x = torch.randn(16, 4096, 512, device=dev).bfloat16()
else:
    vq = vqb.VectorQuantize(dim=256, codebook_size=1024).to(dev)
    x = torch.randn(64, 4096, 256, device=dev).bfloat16()
with torch.no_grad():
    e = torch.randn_like(vq._codebook.embed)
    if which == "4":
        e = torch.nn.functional.normalize(e, dim=-1)
    vq._codebook.embed.copy_(e); vq._codebook.embed_avg.copy_(e)
for _ in range(4):
    q, i, l = vq(x)
torch.cuda.synchronize()
print("ok", float(l))
