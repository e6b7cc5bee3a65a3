"""Driver for ncu captures of the kernels AS THEY RUN in a step (run under gpurun + ncu, VQB_GRAPH=0).

    python scripts/ncu_step.py vq      # BASELINE config 2 training step: vq_assign_kernel with the fused tail, the EMA chain
    python scripts/ncu_step.py rvq     # an 8-stage ResidualVQ step + decode (rvq_accumulate_kernel, decode_kernel)
    python scripts/ncu_step.py gemm    # cuBLAS bf16 8192^3 (the tensor-pipe reference the search kernel is compared with)
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vector_quantize_pytorch_b200 as vqb
dev = torch.device("cuda:0")
torch.manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "vq"
if what == "gemm":
    a = torch.randn(8192, 8192, device=dev).bfloat16(); b = torch.randn(8192, 8192, device=dev).bfloat16()
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
elif what == "vq":
    vq = vqb.VectorQuantize(dim=256, codebook_size=1024).to(dev)
    with torch.no_grad():
        e = torch.randn(1, 1024, 256, device=dev); vq._codebook.embed.copy_(e); vq._codebook.embed_avg.copy_(e)
    x = torch.randn(64, 4096, 256, device=dev).bfloat16()
    vq.train()
    for _ in range(4):
        q, i, l = vq(x)
    torch.cuda.synchronize()
else:
    rvq = vqb.ResidualVQ(dim=256, num_quantizers=8, codebook_size=1024, shared_codebook=True).to(dev)
    with torch.no_grad():
        e = torch.randn(1, 1024, 256, device=dev); rvq.layers[0]._codebook.embed.copy_(e); rvq.layers[0]._codebook.embed_avg.copy_(e)
    x = torch.randn(64, 4096, 256, device=dev).bfloat16()
    rvq.train()
    for _ in range(3):
        q, i, l = rvq(x)
        o = rvq.get_output_from_indices(i)
    torch.cuda.synchronize()
print("done", what)
