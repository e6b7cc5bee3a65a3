"""First-contact probe of the kernels on a real B200 (run under gpurun). Prints diagnostics only."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vector_quantize_pytorch_b200 import ops
from oracle import vq_oracle as O

torch.manual_seed(0)
dev = torch.device("cuda:0")
print("device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def f64_scores(x, c, cosine):
    x64, c64 = x.double(), c.double()
    s = x64 @ c64.T
    if not cosine:
        s = s - 0.5 * (c64 * c64).sum(-1)[None, :]
    return s


def probe(N, D, K, dtype, cosine, tag):
    x = torch.randn(N, D, device=dev)
    c = torch.randn(K, D, device=dev)
    if cosine:
        c = torch.nn.functional.normalize(c, dim=-1)
    x = x.to(dtype)
    cb = ops.prepare_codebook(c, cosine)
    torch.cuda.synchronize()
    # operand checks
    hi = c.bfloat16()
    lo = (c - hi.float()).bfloat16()
    Kpad = ops.padded_codes(K)
    ok_hi = torch.equal(cb.planes[0, :K], hi); ok_lo = torch.equal(cb.planes[1, :K], lo)
    n2 = (c.double() ** 2).sum(-1).float()
    print(f"[{tag}] Kpad={Kpad} planes hi/lo ok: {ok_hi} {ok_lo}  cnorm2 err {(cb.cnorm2 - n2).abs().max().item():.2e} "
          f"cmax {cb.cmax.item():.4f} vs {n2.sqrt().max().item():.4f} bias_pad_inf {bool(torch.isinf(cb.bias[K:]).all())}")
    res = ops.search(x, cb, c, debug_best=True, fix=False)
    torch.cuda.synchronize()
    xe = res.x_eff.float()
    s = f64_scores(xe, c, cosine)
    true_best, true_idx = s.max(-1)
    got = s.gather(1, res.idx.long()[:, None])[:, 0]
    err = (res.best.double() - got).abs()
    scale = xe.double().norm(dim=-1) * c.double().norm(dim=-1).max()
    print(f"[{tag}] raw mismatches vs f64 argmax: {(res.idx.long() != true_idx).sum().item()}/{N}; "
          f"score abs err max {err.max().item():.3e} mean {err.mean().item():.3e}; rel(|x||c|max) max {(err / scale).max().item():.3e} "
          f"(2^-16={2**-16:.3e}); flagged {res.flag_count.item()}")
    res2 = ops.search(x, cb, c)
    torch.cuda.synchronize()
    # oracle
    xo = xe.cpu().numpy()
    co = c.cpu().numpy()
    oi = O.argmax_first(O.scores(xo, co, cosine))
    mism = (res2.idx.cpu().numpy() != oi)
    _, gap = O.top2_gap_f64(xo, co, cosine)
    print(f"[{tag}] after fix: mismatches vs numpy oracle {mism.sum()} (non-tie: {(mism & (gap > 2e-6)).sum()}), flagged {res2.flag_count.item()}")
    # gather + loss
    q = torch.empty_like(x); i64 = torch.empty(N, dtype=torch.int64, device=dev)
    ls = torch.zeros(1, dtype=torch.float64, device=dev)
    ops.gather(res2.x_eff, c, res2.idx, q_out=q, idx64_out=i64, loss_sum=ls)
    torch.cuda.synchronize()
    qref = c[res2.idx.long()].to(dtype)
    lref = ((qref.float() - res2.x_eff.float()) ** 2)
    if dtype == torch.bfloat16:
        lref = lref.bfloat16().float()
    print(f"[{tag}] gather q exact: {torch.equal(q, qref)} idx64 ok: {torch.equal(i64, res2.idx.long())} "
          f"loss_sum rel err {abs(ls.item() - lref.double().sum().item()) / lref.double().sum().item():.2e}")
    # ema stats
    st = ops.ema_stats(res2.x_eff, res2.idx, K)
    torch.cuda.synchronize()
    off = ops.stats_offset(K)
    cs_ref = torch.bincount(res2.idx.long(), minlength=K).float()
    es_ref = torch.zeros(K, D, device=dev, dtype=torch.float64).index_add_(0, res2.idx.long(), res2.x_eff.double())
    es = st[off:].view(K, D)
    print(f"[{tag}] stats cluster_size ok: {torch.equal(st[:K], cs_ref)}  embed_sum max abs err {(es.double() - es_ref).abs().max().item():.3e} "
          f"(max |sum| {es_ref.abs().max().item():.1f})")
    # ema apply vs oracle
    state = O.CodebookState.from_embed(co)
    cs = torch.ones(K, device=dev); ea = c.clone(); emb = c.clone()
    ops.ema_apply(cs, ea, emb, st, cb, decay=0.8, eps=1e-5, do_lerp=True, do_normalise=True)
    torch.cuda.synchronize()
    O.track_stats(state, xo, res2.idx.cpu().numpy().astype(np.int64), 0.8)
    O.update_ema(state, 1e-5, cosine)
    print(f"[{tag}] ema: cluster_size err {np.abs(cs.cpu().numpy() - state.cluster_size).max():.2e} embed_avg err "
          f"{np.abs(ea.cpu().numpy() - state.embed_avg).max():.2e} embed err {np.abs(emb.cpu().numpy() - state.embed).max():.2e}")
    hi2 = emb.bfloat16()
    print(f"[{tag}] refreshed planes ok: {torch.equal(cb.planes[0, :K], hi2)}")


def timeit(N, D, K, dtype, cosine=False, iters=10):
    x = torch.randn(N, D, device=dev).to(dtype)
    c = torch.randn(K, D, device=dev)
    cb = ops.prepare_codebook(c, cosine)
    for _ in range(3):
        r = ops.search(x, cb, c)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        r = ops.search(x, cb, c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"[time] search N={N} D={D} K={K} {dtype}: {ms:.3f} ms  -> {N / ms * 1e3:.3e} vec/s, {2 * N * K * D / ms / 1e9:.1f} TFLOP/s algorithmic; flagged {r.flag_count.item()}")


if __name__ == "__main__":
    probe(1000, 256, 1024, torch.bfloat16, False, "bf16-euclid")
    probe(777, 64, 96, torch.float32, False, "fp32-euclid-small")
    probe(1000, 256, 512, torch.float32, False, "fp32-euclid")
    probe(515, 128, 40, torch.bfloat16, True, "bf16-cosine-smallK")
    probe(2048, 512, 2048, torch.bfloat16, True, "bf16-cosine-D512")
    probe(300, 32, 5, torch.float32, False, "tiny")
    timeit(262144, 256, 1024, torch.bfloat16)
    timeit(262144, 256, 1024, torch.float32)
    timeit(65536, 512, 16384, torch.bfloat16, cosine=True, iters=3)
