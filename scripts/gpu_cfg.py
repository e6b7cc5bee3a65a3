"""A few forwards of one BASELINE.json config (for `ncu` launch lists): gpu_cfg.py {2|3|4}"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vector_quantize_pytorch_b200 as vqb
dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "4"
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
