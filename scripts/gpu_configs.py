"""Module-level timings of every BASELINE.json config on one B200 (diagnostic; numbers go to profiles/)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vector_quantize_pytorch_b200 as vqb

dev = torch.device("cuda:0")


def warm(mod, cosine=False):
    for m in mod.modules():
        if isinstance(m, vqb.Codebook):
            e = torch.randn_like(m.embed)
            if cosine:
                e = torch.nn.functional.normalize(e, dim=-1)
            m.embed.copy_(e); m.embed_avg.copy_(e)


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(1234)
    out = []
    cases = [
        ("cfg1 VQ D=256 K=512 x=(1,1024,256) fp32", lambda: vqb.VectorQuantize(dim=256, codebook_size=512), (1, 1024, 256), torch.float32, False, 1),
        ("cfg2 VQ D=256 K=1024 x=(64,4096,256) bf16", lambda: vqb.VectorQuantize(dim=256, codebook_size=1024), (64, 4096, 256), torch.bfloat16, False, 1),
        ("cfg2f VQ D=256 K=1024 x=(64,4096,256) fp32", lambda: vqb.VectorQuantize(dim=256, codebook_size=1024), (64, 4096, 256), torch.float32, False, 1),
        ("cfg3 RVQ Q=8 shared K=1024 x=(32,8192,256) fp32", lambda: vqb.ResidualVQ(dim=256, num_quantizers=8, codebook_size=1024, shared_codebook=True), (32, 8192, 256), torch.float32, False, 8),
        ("cfg3b RVQ Q=8 shared K=1024 x=(32,8192,256) bf16", lambda: vqb.ResidualVQ(dim=256, num_quantizers=8, codebook_size=1024, shared_codebook=True), (32, 8192, 256), torch.bfloat16, False, 8),
        ("cfg4 VQ cosine D=512 K=16384 x=(16,4096,512) bf16", lambda: vqb.VectorQuantize(dim=512, codebook_size=16384, use_cosine_sim=True), (16, 4096, 512), torch.bfloat16, True, 1),
        ("cfg4f VQ cosine D=512 K=16384 x=(16,4096,512) fp32", lambda: vqb.VectorQuantize(dim=512, codebook_size=16384, use_cosine_sim=True), (16, 4096, 512), torch.float32, True, 1),
        ("cfg5 GRVQ G=2 Q=8 K=1024 shard x=(8,4096,256) fp32", lambda: vqb.GroupedResidualVQ(dim=256, groups=2, num_quantizers=8, codebook_size=1024), (8, 4096, 256), torch.float32, False, 16),
    ]
    for name, build, shape, dt, cosine, stages in cases:
        mod = build().to(dev)
        with torch.no_grad():
            warm(mod, cosine)
        mod.train()
        x = torch.randn(*shape, device=dev).to(dt)
        ms = timeit(lambda: mod(x), iters=5 if stages > 1 else 10)
        n = shape[0] * shape[1]
        rec = dict(config=name, ms=ms, vectors_per_s=n / ms * 1e3, stage_vectors_per_s=n * stages / ms * 1e3)
        out.append(rec)
        print(json.dumps(rec), flush=True)
    return out


if __name__ == "__main__":
    main()
