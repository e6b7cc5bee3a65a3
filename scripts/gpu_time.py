"""Per-op CUDA-event timings on a B200 (diagnostic; run under gpurun)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vector_quantize_pytorch_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(N, D, K, dtype, cosine=False, n_passes=0):
    x = torch.randn(N, D, device=dev).to(dtype)
    c = torch.randn(K, D, device=dev)
    if cosine:
        c = torch.nn.functional.normalize(c, dim=-1)
    cb = ops.prepare_codebook(c, cosine)
    # spin the clocks up
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(20):
        a @ a
    torch.cuda.synchronize()
    t_nofix = ev_time(lambda: ops.search(x, cb, c, fix=False, n_passes=n_passes))
    t_fix = ev_time(lambda: ops.search(x, cb, c, n_passes=n_passes))
    r = ops.search(x, cb, c, n_passes=n_passes)
    cnt = r.flag_count.item()
    over = int((r.flagged[:cnt, 1] > 2).sum().item()) if cnt else 0
    q = torch.empty_like(x); i64 = torch.empty(N, dtype=torch.int64, device=dev); ls = torch.zeros(1, dtype=torch.float64, device=dev)
    t_gather = ev_time(lambda: ops.gather(r.x_eff, c, r.idx, q_out=q, idx64_out=i64, loss_sum=ls))
    t_stats = ev_time(lambda: ops.ema_stats(r.x_eff, r.idx, K))
    st = ops.ema_stats(r.x_eff, r.idx, K)
    cs = torch.ones(K, device=dev); ea = c.clone(); emb = c.clone()
    t_apply = ev_time(lambda: ops.ema_apply(cs, ea, emb, st, cb, decay=0.8, eps=1e-5, do_lerp=True, do_normalise=True))
    print(f"N={N} D={D} K={K} {str(dtype)[6:]} cos={cosine} passes={n_passes}: search(nofix) {t_nofix*1e3:.0f}us  search {t_fix*1e3:.0f}us  "
          f"gather {t_gather*1e3:.0f}us  stats {t_stats*1e3:.0f}us  apply {t_apply*1e3:.0f}us  flagged {cnt} (overflow {over})  "
          f"=> search-only {N/t_nofix*1e3:.3e} vec/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "cfg2"):
        run(262144, 256, 1024, torch.bfloat16)
        run(262144, 256, 1024, torch.bfloat16, n_passes=1)
        run(262144, 256, 1024, torch.float32)
    if which in ("all", "cfg4"):
        run(65536, 512, 16384, torch.bfloat16, cosine=True)
    if which in ("all", "cfg5"):
        run(32768, 128, 1024, torch.float32)
