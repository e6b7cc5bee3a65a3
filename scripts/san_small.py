"""Small shapes through every kernel of the step (both dtypes, both metrics, streamed A, ResidualVQ program + decode) for
compute-sanitizer:  VQB_GRAPH=0 compute-sanitizer --tool {memcheck|racecheck|synccheck} python scripts/san_small.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vector_quantize_pytorch_b200 as vqb
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, T, D, K, dt, cos) in ((1, 300, 32, 64, torch.float32, False), (2, 1024, 256, 1024, torch.bfloat16, False),
                              (1, 512, 512, 600, torch.float32, True), (1, 700, 128, 1000, torch.bfloat16, True)):
    vq = vqb.VectorQuantize(dim=D, codebook_size=K, use_cosine_sim=cos).to(dev)
    x = torch.randn(B, T, D, device=dev).to(dt)
    for _ in range(2):
        q, i, l = vq(x)
    torch.cuda.synchronize()
    print("ok vq", B, T, D, K, dt, cos)
rvq = vqb.ResidualVQ(dim=64, num_quantizers=3, codebook_size=96, shared_codebook=True).to(dev)
y = torch.randn(2, 2400, 64, device=dev).bfloat16()
for _ in range(2):
    q, i, l = rvq(y)
o = rvq.get_output_from_indices(i); torch.cuda.synchronize(); print("ok rvq")
