"""CPU oracle for the nearest-code search / gather / EMA hot path  —  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm of lucidrains/vector-quantize-pytorch v1.31.0
for the ONE path this repo accelerates.  It is imported only by `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py`.  The product package
(`vector_quantize_pytorch_b200`) never imports it and has no CPU fallback.

Parity pin: the reference ships NO golden vectors for this path, so the oracle is
pinned against outputs of the reference itself: `oracle/gen_golden.py` imports the unmodified
reference from /root/reference (with the `einx` shim), runs it on seeded inputs and commits
(inputs, initial buffers, outputs, post-step buffers) under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against every fixture.

Citations are `file:line` into /root/reference/vector_quantize_pytorch/ :
  vqp = vector_quantize_pytorch.py , rvq = residual_vq.py

Arithmetic notes (all arrays are float32 unless stated):
  * `Codebook.forward` upcasts its input to fp32 (vqp:692) and the codebook buffers are fp32, so the
    search is an fp32 computation for every input dtype.
  * bf16 tensors are represented as float32 arrays holding bf16-representable values; `bf16_round`
    applies round-to-nearest-even exactly where the reference's tensors are bf16.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------------
# dtype helpers
# --------------------------------------------------------------------------------------------


def bf16_round(a: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16 -> float32 (what `tensor.bfloat16().float()` does)."""
    a = np.ascontiguousarray(a, dtype=F32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(F32).copy()
    nan = np.isnan(a)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(a.shape)


def cast_like(a: np.ndarray, dtype: str) -> np.ndarray:
    """`tensor.type(dtype)` for dtype in {'fp32','bf16'} (vqp:1178)."""
    if dtype == "fp32":
        return np.asarray(a, dtype=F32)
    if dtype == "bf16":
        return bf16_round(a)
    raise ValueError(dtype)


def l2norm(t: np.ndarray, dtype: str = "fp32", eps: float = 1e-6) -> np.ndarray:
    """vqp:37-38  F.normalize(t, p=2, dim=-1, eps): t / max(||t||, eps), evaluated in t's dtype.

    For bf16 the norm is accumulated in fp32 and rounded to bf16, and the quotient is rounded to
    bf16 (verified against torch 2.11 CPU: 0 mismatches over 2M elements).
    """
    t = np.asarray(t, dtype=F32)
    n = np.sqrt(np.sum(t * t, axis=-1, keepdims=True, dtype=F32), dtype=F32)
    if dtype == "bf16":
        n = bf16_round(n)
    n = np.maximum(n, F32(eps))
    out = (t / n).astype(F32)
    return bf16_round(out) if dtype == "bf16" else out


# --------------------------------------------------------------------------------------------
# Codebook state  (vqp:415-423)
# --------------------------------------------------------------------------------------------


@dataclass
class CodebookState:
    """Persistent buffers of `Codebook` for num_codebooks == 1 (leading H dim squeezed)."""

    embed: np.ndarray  # (K, D) fp32           vqp:419-423
    embed_avg: np.ndarray  # (K, D) fp32       vqp:417  (init: clone of embed)
    cluster_size: np.ndarray  # (K,) fp32      vqp:416  (init: ones)
    initted: bool = True  # vqp:415

    @staticmethod
    def from_embed(embed: np.ndarray) -> "CodebookState":
        embed = np.array(embed, dtype=F32)
        return CodebookState(embed.copy(), embed.copy(), np.ones(embed.shape[0], dtype=F32))

    def copy(self) -> "CodebookState":
        return CodebookState(self.embed.copy(), self.embed_avg.copy(), self.cluster_size.copy(), self.initted)


# --------------------------------------------------------------------------------------------
# distance + argmax   (vqp:58-62, :741-747, :130-145)
# --------------------------------------------------------------------------------------------


def neg_cdist(x: np.ndarray, y: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """-cdist(x, y) (vqp:58-62 negated at vqp:743), evaluated in the reference's operation order."""
    x2 = np.sum(x * x, axis=-1, dtype=F32)  # vqp:59
    y2 = np.sum(y * y, axis=-1, dtype=F32)  # vqp:60
    xy = (x @ y.T).astype(F32) * F32(-2.0)  # vqp:61
    d2 = (x2[:, None] + y2[None, :]) + xy  # vqp:62  (left-to-right)
    return -np.sqrt(np.maximum(d2, F32(eps)), dtype=F32)


def scores(x: np.ndarray, embed: np.ndarray, cosine: bool) -> np.ndarray:
    """`dist` of vqp:741 (cosine: x·cᵀ) / vqp:743 (euclid: -cdist).  (N, K) fp32."""
    if cosine:
        return (x @ embed.T).astype(F32)
    return neg_cdist(x, embed)


def argmax_first(dist: np.ndarray) -> np.ndarray:
    """vqp:140 `argmax(dim=-1)`; torch and numpy both return the FIRST maximal index."""
    return np.argmax(dist, axis=-1).astype(np.int64)


# --------------------------------------------------------------------------------------------
# EMA   (vqp:76-97, :152-154, :576-617)
# --------------------------------------------------------------------------------------------


def batch_stats(x: np.ndarray, ind: np.ndarray, K: int, faithful: bool = False):
    """cluster_size = onehot.sum(1) (vqp:602), embed_sum = xᵀ·onehot (vqp:605).

    faithful=True materialises the one-hot and does the GEMM like the reference; the default uses
    bincount / add.at, which is the same sum in a different order (1.7e-7 rel).
    """
    if faithful:
        onehot = np.zeros((x.shape[0], K), dtype=F32)
        onehot[np.arange(x.shape[0]), ind] = 1.0
        return onehot.sum(axis=0, dtype=F32), (onehot.T @ x).astype(F32)
    cs = np.bincount(ind, minlength=K).astype(F32)
    es = np.zeros((K, x.shape[1]), dtype=np.float64)
    np.add.at(es, ind, x.astype(np.float64))
    return cs, es.astype(F32)


def ema_inplace(old: np.ndarray, new: np.ndarray, decay: float, weight=None) -> None:
    """vqp:76-97  old.lerp_(new, (1-decay)*weight)  (torch lerp: old + w*(new-old), fp32)."""
    w = F32(1.0 - decay)
    if weight is not None:
        weight = np.asarray(weight, dtype=F32)
        w = (w * weight).astype(F32)
        if old.ndim == 2:
            w = w[:, None]
    # torch.lerp for weight < 0.5 : start + weight * (end - start)
    wb = np.broadcast_to(np.asarray(w, dtype=F32), old.shape)
    lo = old + wb * (new - old)
    hi = new - (new - old) * (F32(1.0) - wb)
    old[...] = np.where(wb < F32(0.5), lo, hi).astype(F32)


def laplace_smoothing(x: np.ndarray, n_categories: int, eps: float = 1e-5) -> np.ndarray:
    """vqp:152-154."""
    denom = np.sum(x, axis=-1, keepdims=True, dtype=F32)
    return (x + F32(eps)) / (denom + F32(n_categories * eps))


def update_ema(state: CodebookState, eps: float, cosine: bool) -> None:
    """vqp:576-584: embed = embed_avg / (laplace(cluster_size) * cluster_size.sum()); l2norm if cosine."""
    K = state.cluster_size.shape[0]
    cs = laplace_smoothing(state.cluster_size, K, eps) * np.sum(state.cluster_size, dtype=F32)
    embed = (state.embed_avg / cs[:, None]).astype(F32)
    if cosine:
        embed = l2norm(embed)
    state.embed[...] = embed


def track_stats(state: CodebookState, x: np.ndarray, ind: np.ndarray, decay: float, all_reduce=None,
                ema_update_weight=None, faithful: bool = False) -> None:
    """vqp:586-617 for the non-accumulating branch."""
    K = state.cluster_size.shape[0]
    cs, es = batch_stats(x, ind, K, faithful)
    if all_reduce is not None:  # vqp:603, :607 (SUM over ranks)
        cs = all_reduce(cs)
        es = all_reduce(es)
    ema_inplace(state.cluster_size, cs, decay, ema_update_weight)  # vqp:616
    ema_inplace(state.embed_avg, es, decay, ema_update_weight)  # vqp:617


# --------------------------------------------------------------------------------------------
# Codebook.forward  (vqp:674-791)
# --------------------------------------------------------------------------------------------


def expire_codes(state: CodebookState, samples: np.ndarray, threshold: float, reset: float, pick_fn, cosine: bool = False,
                 dtype: str = "fp32") -> int:
    """vqp:564-574 + :544-562 (`replace`) + :156-163 (`sample_vectors`): codes whose EMA cluster size fell below
    `threshold` are re-seeded from `samples` (N, D).  `pick_fn(n, num)` supplies the sampled row ids — the reference
    draws torch.randperm(n)[:num] (n >= num) or torch.randint(0, n, (num,)); the caller replays the same RNG.
    Returns the number of replaced codes."""
    expired = state.cluster_size < F32(threshold)  # vqp:568
    num = int(expired.sum())
    if num == 0:  # vqp:570
        return 0
    samples = np.asarray(samples, dtype=F32).reshape(-1, samples.shape[-1])
    if cosine:  # vqp:545-546, in the dtype the samples arrive in (fp32 from Codebook.forward, x.dtype from rvq:601)
        samples = l2norm(samples, dtype)
    picks = np.asarray(pick_fn(samples.shape[0], num)).astype(np.int64)
    sampled = samples[picks]
    state.embed[expired] = sampled  # vqp:560
    state.cluster_size[expired] = F32(reset)  # vqp:561
    state.embed_avg[expired] = sampled * F32(reset)  # vqp:562
    return num


def kmeans(samples: np.ndarray, num_clusters: int, num_iters: int, cosine: bool, pick_fn):
    """vqp:238-278 for one codebook: Lloyd iterations from `sample_vectors(samples, K)` (vqp:156-163; `pick_fn(n, K)`
    supplies the sampled row ids).  Empty clusters keep their previous mean (vqp:271-275).  Returns (means, bins)."""
    samples = np.asarray(samples, dtype=F32)
    means = samples[np.asarray(pick_fn(samples.shape[0], num_clusters)).astype(np.int64)].copy()
    bins = np.zeros(num_clusters, dtype=np.int64)
    for _ in range(num_iters):
        dists = (samples @ means.T).astype(F32) if cosine else neg_cdist(samples, means)  # vqp:251-254
        buckets = argmax_first(dists)  # vqp:256
        bins = np.bincount(buckets, minlength=num_clusters)  # vqp:257
        zero = bins == 0
        new = np.zeros((num_clusters, samples.shape[1]), dtype=np.float64)
        np.add.at(new, buckets, samples.astype(np.float64))  # vqp:265
        new = (new / np.maximum(bins, 1)[:, None]).astype(F32)  # vqp:266
        if cosine:
            new = l2norm(new)  # vqp:269-270
        means = np.where(zero[:, None], means, new)  # vqp:272-276
    return means.astype(F32), bins


def init_embed(state: CodebookState, data: np.ndarray, *, kmeans_iters: int, cosine: bool, eps: float, pick_fn) -> None:
    """vqp:451-473: k-means initialisation on the first batch, then update_ema."""
    if state.initted:
        return
    K = state.cluster_size.shape[0]
    embed, bins = kmeans(data, K, kmeans_iters, cosine, pick_fn)
    cluster_size = bins.astype(F32)
    state.embed_avg[...] = embed * cluster_size[:, None]  # vqp:467-469
    state.cluster_size[...] = cluster_size  # vqp:470
    update_ema(state, eps, cosine)  # vqp:471
    state.initted = True


def codebook_forward(x: np.ndarray, state: CodebookState, *, cosine: bool = False, training: bool = True,
                     decay: float = 0.8, eps: float = 1e-5, ema_update: bool = True,
                     manual_ema_update: bool = False, freeze_codebook: bool = False, all_reduce=None,
                     ema_update_weight=None, faithful: bool = False, threshold_ema_dead_code: float = 0,
                     pick_fn=None, accum: dict | None = None, accum_ema_update: bool = False, kmeans_iters: int = 10,
                     row_mask: np.ndarray | None = None):
    """x: (N, D) fp32 (already upcast, vqp:692; already l2-normalised if cosine, vqp:1159).
    row_mask (N,) bool (vqp:700-701): EVERY row is searched, but only the rows with mask True enter the k-means init
    (vqp:456-458), the batch statistics (vqp:599-600: their one-hot rows are zeroed) and the expiry samples (vqp:549-550).

    Returns (quantize (N, D) fp32, embed_ind (N,) int64).  Mutates `state` like the reference:
    the search and the returned quantize use the PRE-update codebook (vqp:766 precedes :783-784).
    """
    x = np.asarray(x, dtype=F32)
    xs = x if row_mask is None else x[row_mask]   # the rows that count for init / statistics / expiry
    if not state.initted:  # vqp:703
        init_embed(state, xs, kmeans_iters=kmeans_iters, cosine=cosine, eps=eps, pick_fn=pick_fn)
    embed = state.embed  # vqp:710-712
    dist = scores(x, embed, cosine)  # vqp:741 / :743
    ind = argmax_first(dist)  # vqp:747 -> :140
    if faithful and training:  # vqp:766  onehot @ embed  (exact row copy)
        onehot = np.zeros((x.shape[0], embed.shape[0]), dtype=F32)
        onehot[np.arange(x.shape[0]), ind] = 1.0
        quantize = (onehot @ embed).astype(F32)
    else:  # vqp:779-781 gather; bit-identical to the one-hot product
        quantize = embed[ind].copy()
    has_expiry = threshold_ema_dead_code > 0
    if training and not freeze_codebook and (ema_update or has_expiry):  # vqp:783-784, :619-641
        ind_all = ind
        if row_mask is not None:
            x, ind = xs, ind[row_mask]
        if accum is not None:  # vqp:70-74, :80-82, :612-614: statistics parked on `.grad` across calls
            K = state.cluster_size.shape[0]
            cs, es = batch_stats(x, ind, K, faithful)
            if callable(ema_update_weight):
                ema_update_weight = ema_update_weight(es, cs)
            if accum_ema_update:
                accum["cs"] = accum.get("cs", 0) + cs
                accum["es"] = accum.get("es", 0) + es
                return quantize, ind_all
            cs = cs + accum.pop("cs", 0)
            es = es + accum.pop("es", 0)
            ema_inplace(state.cluster_size, cs, decay, ema_update_weight)
            ema_inplace(state.embed_avg, es, decay, ema_update_weight)
        else:
            if callable(ema_update_weight):
                K = state.cluster_size.shape[0]
                cs, es = batch_stats(x, ind, K, faithful)
                ema_update_weight = ema_update_weight(es, cs)
            track_stats(state, x, ind, decay, all_reduce, ema_update_weight, faithful)
        if ema_update and not manual_ema_update:  # vqp:638-639
            update_ema(state, eps, cosine)
        if has_expiry:  # vqp:641
            expire_codes(state, x, threshold_ema_dead_code, threshold_ema_dead_code, pick_fn, cosine)
        ind = ind_all
    return quantize, ind


# --------------------------------------------------------------------------------------------
# VectorQuantize.forward  (vqp:1093-1403) — default path: heads=1, no projection, no mask
# --------------------------------------------------------------------------------------------


def mse_loss(a: np.ndarray, b: np.ndarray, dtype: str):
    """F.mse_loss(a, b) (vqp:1327) in the tensors' dtype.

    bf16: torch rounds (a-b)^2 to bf16 per element, reduces in fp32, rounds the mean to bf16.
    Returns (value_as_reference_returns_it, fp32_mean_before_final_rounding).
    """
    d = (a.astype(F32) - b.astype(F32))
    sq = (d * d).astype(F32)
    if dtype == "bf16":
        sq = bf16_round(sq)
    mean = F32(np.sum(sq, dtype=np.float64) / sq.size)
    if dtype == "bf16":
        return F32(bf16_round(np.array([mean], dtype=F32))[0]), mean
    return mean, mean


@dataclass
class VQConfig:
    dim: int
    codebook_size: int
    use_cosine_sim: bool = False
    decay: float = 0.8  # vqp:810
    eps: float = 1e-5  # vqp:811
    commitment_weight: float = 1.0  # vqp:822
    manual_ema_update: bool = False
    ema_update: bool = True
    threshold_ema_dead_code: float = 0  # vqp:818
    kmeans_iters: int = 10  # vqp:816
    heads: int = 1  # vqp:807 (codebook shared across the heads; `dim` is then the per-head codebook dim)
    separate_codebook_per_head: bool = False  # vqp:808: `state` is then a list of `heads` CodebookState


def vq_forward(x: np.ndarray, dtype: str, state: CodebookState, cfg: VQConfig, *, training: bool = True,
               freeze_codebook: bool = False, all_reduce=None, faithful: bool = False, ema_update_weight=None,
               pick_fn=None, accum: dict | None = None, accum_ema_update: bool = False, mask: np.ndarray | None = None,
               return_zeros_for_masked_padding: bool = True):
    """x: (..., D) values of dtype `dtype` held in float32.  x.requires_grad is False (bench setting).
    mask (B, N) bool — variable-length input (vqp:1116-1119; heads == 1 here): see the masked branches below.

    Returns (quantize (..., D) in dtype, indices (...,) int64, loss fp32 scalar, loss_fp32_unrounded).
    """
    shape = x.shape
    if mask is not None:
        assert cfg.heads == 1 and mask.shape == tuple(shape[:-1])
    if cfg.heads > 1 and cfg.separate_codebook_per_head:
        # vqp:1044-1049 'b n (h d) -> h b n d', Codebook(num_codebooks = h): h independent codebooks, processed in head order
        # (the per-head RNG draws of k-means / expiry follow that order, vqp:166-167); ONE mse over all heads (vqp:1327)
        b, n, hd = shape
        h, d = cfg.heads, hd // cfg.heads
        sub = VQConfig(**{**cfg.__dict__, "heads": 1, "separate_codebook_per_head": False})
        xc = cast_like(x, dtype)
        qs, inds, l32 = [], [], []
        for i in range(h):
            q, ind, _, loss32 = vq_forward(xc[..., i * d:(i + 1) * d], dtype, state[i], sub, training=training,
                                           freeze_codebook=freeze_codebook, all_reduce=all_reduce, faithful=faithful, pick_fn=pick_fn)
            qs.append(q); inds.append(ind); l32.append(loss32)
        q = np.concatenate(qs, axis=-1)
        ind = np.stack(inds, axis=-1)
        loss32 = F32(np.mean(np.array(l32, dtype=F32), dtype=F32))
        loss = loss32
        if dtype == "bf16" and training and cfg.commitment_weight > 0:
            loss = bf16_round(np.array([loss32], dtype=F32))[0]
        return q, ind, F32(loss), loss32
    if cfg.heads > 1:  # vqp:1044-1049: 'b n (h d) -> 1 (b h) n d' — every head's sub-vector is a row of the ONE codebook
        b, n, hd = shape
        h, d = cfg.heads, hd // cfg.heads
        xs = x.reshape(b, n, h, d).transpose(0, 2, 1, 3).reshape(b * h, n, d)
        q, ind, loss, loss32 = vq_forward(xs, dtype, state, VQConfig(**{**cfg.__dict__, "heads": 1}), training=training,
                                          freeze_codebook=freeze_codebook, all_reduce=all_reduce, faithful=faithful,
                                          ema_update_weight=ema_update_weight, pick_fn=pick_fn, accum=accum,
                                          accum_ema_update=accum_ema_update)
        q = q.reshape(b, h, n, d).transpose(0, 2, 1, 3).reshape(b, n, hd)          # vqp:1354-1358
        ind = ind.reshape(b, h, n).transpose(0, 2, 1)                              # vqp:1266-1270: 'b n h'
        return q, ind, loss, loss32
    x = cast_like(x, dtype).reshape(-1, shape[-1])
    orig_input = x  # vqp:1108
    row_mask = None if mask is None else np.asarray(mask, dtype=bool).reshape(-1)
    if cfg.use_cosine_sim:  # vqp:1159 -> :376 : l2norm in the INPUT dtype
        x = l2norm(x, dtype)
    quantize, ind = codebook_forward(  # vqp:1176
        x, state, cosine=cfg.use_cosine_sim, training=training, decay=cfg.decay, eps=cfg.eps,
        ema_update=cfg.ema_update, manual_ema_update=cfg.manual_ema_update, freeze_codebook=freeze_codebook,
        all_reduce=all_reduce, faithful=faithful, ema_update_weight=ema_update_weight,
        threshold_ema_dead_code=cfg.threshold_ema_dead_code, pick_fn=pick_fn, accum=accum, accum_ema_update=accum_ema_update,
        kmeans_iters=cfg.kmeans_iters, row_mask=row_mask)
    quantize = cast_like(quantize, dtype)  # vqp:1178
    loss = F32(0.0)
    loss_f32 = F32(0.0)
    if training and cfg.commitment_weight > 0:  # vqp:1282-1329
        if row_mask is not None:  # vqp:1317-1325: mse(reduction none) against the ORIGINAL input, mean over the unmasked elements
            cl, cl32 = mse_loss(quantize[row_mask], orig_input[row_mask], dtype)
        else:
            cl, cl32 = mse_loss(quantize, x, dtype)  # vqp:1327 (x = post-l2norm input)
        prod = cl * F32(cfg.commitment_weight)  # vqp:1329: a bf16 tensor times a python float stays bf16
        if dtype == "bf16":
            prod = bf16_round(np.array([prod], dtype=F32))[0]
        loss = F32(F32(0.0) + prod)  # vqp:1282: promoted by the fp32 accumulator
        loss_f32 = F32(cl32 * F32(cfg.commitment_weight))
    if row_mask is not None:  # vqp:1378-1396: padding comes back as zeros (or the input) and index -1
        fill = np.zeros_like(orig_input) if return_zeros_for_masked_padding else orig_input
        quantize = np.where(row_mask[:, None], quantize, fill)
        ind = np.where(row_mask, ind, -1)
    return quantize.reshape(shape), ind.reshape(shape[:-1]), loss, loss_f32


def vq_forward_layout(x: np.ndarray, dtype: str, state, cfg: VQConfig, *, layout: str | None, **kw):
    """VectorQuantize.forward on the reference's other input layouts (everything else as `vq_forward`):
      "single"        x (b, d): one token per batch element, 'b d -> b 1 d' and back (vqp:1121-1125, :1277, :1375-1376)
      "image"         x (b, c, h, w)    'b c h w -> b (h w) c'      (vqp:1136-1139); indices come back (b, h, w[, heads])
      "3d"            x (b, c, d, h, w) 'b c d h w -> b (d h w) c'  (vqp:1141-1144); indices (b, d, h, w[, heads])
      "channel_first" x (b, d, n)       'b d n -> b n d'            (vqp:1146-1147); indices stay (b, n)
    quantize is restored to the input layout (vqp:1364-1376)."""
    if layout is None:
        return vq_forward(x, dtype, state, cfg, **kw)
    if layout == "single":
        q, ind, loss, l32 = vq_forward(x[:, None, :], dtype, state, cfg, **kw)
        return q[:, 0], ind[:, 0], loss, l32
    if layout == "channel_first":
        q, ind, loss, l32 = vq_forward(np.ascontiguousarray(x.transpose(0, 2, 1)), dtype, state, cfg, **kw)
        return np.ascontiguousarray(q.transpose(0, 2, 1)), ind, loss, l32
    assert layout in ("image", "3d")
    b, c, *sp = x.shape
    rows = np.ascontiguousarray(np.moveaxis(x, 1, -1)).reshape(b, -1, c)
    q, ind, loss, l32 = vq_forward(rows, dtype, state, cfg, **kw)
    q = np.ascontiguousarray(np.moveaxis(q.reshape(b, *sp, c), -1, 1))
    return q, ind.reshape(b, *sp, *ind.shape[2:]), loss, l32


# --------------------------------------------------------------------------------------------
# ResidualVQ.forward  (rvq:384-630) — plain loop: no beam, no dropout, no projection
# --------------------------------------------------------------------------------------------


def _bin(a: np.ndarray, dtype: str) -> np.ndarray:
    return bf16_round(a) if dtype == "bf16" else a.astype(F32)


def rvq_forward(x: np.ndarray, dtype: str, states: list, cfg: VQConfig, *, shared_codebook: bool = False,
                training: bool = True, freeze_codebook: bool = False, all_reduce=None, faithful: bool = False, pick_fn=None,
                mask: np.ndarray | None = None, dropout_index: int | None = None):
    """states: list of Q CodebookState (for shared_codebook all entries are THE SAME object, rvq:302-306).

    dropout_index (training with quantize_dropout, rvq:423-439: `quantize_dropout_index(seed, ...)` below): the layers after it
    are skipped — their indices come back as -1, their losses as 0 (rvq:473-476).

    Returns (quantized_out (..., D) in dtype, indices (..., Q) int64, losses (Q,) fp32, losses_fp32 (Q,)).
    """
    Q = len(states)
    if shared_codebook:
        assert all(s is states[0] for s in states)
    layer_cfg = VQConfig(**{**cfg.__dict__, "manual_ema_update": shared_codebook or cfg.manual_ema_update})  # rvq:213-217
    x = cast_like(x, dtype)
    quantized_out = np.zeros_like(x)  # rvq:410
    residual = x  # rvq:411
    all_ind, all_loss, all_loss32, all_residuals = [], [], [], []
    for q in range(Q):  # rvq:469
        if training and dropout_index is not None and q > dropout_index:  # rvq:473-476
            all_ind.append(np.full(x.shape[:-1], -1, dtype=np.int64))
            all_loss.append(F32(0.0))
            all_loss32.append(F32(0.0))
            continue
        all_residuals.append(residual)  # rvq:489
        quantized, ind, loss, loss32 = vq_forward(residual, dtype, states[q], layer_cfg, training=training,
                                                  freeze_codebook=freeze_codebook, all_reduce=all_reduce,
                                                  faithful=faithful, pick_fn=pick_fn, mask=mask)  # rvq:493 (every layer gets the mask, :495)
        residual = _bin(residual - quantized, dtype)  # rvq:524 (quant_grad_frac=0 -> detach)
        quantized_out = _bin(quantized_out + quantized, dtype)  # rvq:525
        all_ind.append(ind)
        all_loss.append(loss)
        all_loss32.append(loss32)
    if training and shared_codebook:  # rvq:593-601 (not gated by freeze_codebook in the reference either)
        if cfg.ema_update:
            update_ema(states[0], cfg.eps, cfg.use_cosine_sim)
        if cfg.threshold_ema_dead_code > 0:
            # rvq:599-601 -> vqp:1051-1054 -> :573-574: all_residuals '(b) n l d -> b (n l) d' reaches Codebook.expire_codes_
            # whose 'h ... d -> h (...) d' reads the BATCH axis as the codebook axis; `replace` zips it with the (1, K)
            # mask, so only batch element 0's rows are sampled from.  Reproduced as is.
            allr = np.stack([r[0].reshape(-1, r.shape[-1]) for r in all_residuals], axis=1).reshape(-1, x.shape[-1])
            if cfg.use_cosine_sim:
                allr = l2norm(allr, dtype)
            expire_codes(states[0], allr, cfg.threshold_ema_dead_code, cfg.threshold_ema_dead_code, pick_fn, cfg.use_cosine_sim,
                         dtype)
    return (quantized_out, np.stack(all_ind, axis=-1), np.array(all_loss, dtype=F32),
            np.array(all_loss32, dtype=F32))


def quantize_dropout_index(seed: int, num_quantizers: int, cutoff_index: int = 0, multiple_of: int = 1) -> int:
    """rvq:434-439: the last layer that still quantizes, drawn from python's `random.Random(seed)`."""
    import math
    import random
    idx = random.Random(seed).randrange(cutoff_index, num_quantizers)
    if multiple_of != 1:
        idx = math.ceil((idx + 1) / multiple_of) * multiple_of - 1  # rvq:39-40 round_up_multiple
    return idx


def grouped_rvq_forward(x: np.ndarray, dtype: str, group_states: list, cfg: VQConfig, **kw):
    """rvq:676-724.  cfg.dim is the per-group dim; x: (..., G*dim).

    Returns (quantized (..., G*dim), indices (G, ..., Q), losses (G, Q), losses_fp32 (G, Q)).
    """
    G = len(group_states)
    chunks = np.split(np.asarray(x, dtype=F32), G, axis=-1)  # rvq:690
    outs = [rvq_forward(c, dtype, st, cfg, **kw) for c, st in zip(chunks, group_states)]  # rvq:706
    return (np.concatenate([o[0] for o in outs], axis=-1), np.stack([o[1] for o in outs]),  # rvq:719-721
            np.stack([o[2] for o in outs]), np.stack([o[3] for o in outs]))


# --------------------------------------------------------------------------------------------
# decode  (vqp:998-1022, rvq:324-382)
# --------------------------------------------------------------------------------------------


def vq_codes_from_indices(embed: np.ndarray, indices: np.ndarray) -> np.ndarray:
    return embed[indices]  # vqp:1003


def rvq_output_from_indices(embeds: list, indices: np.ndarray) -> np.ndarray:
    """rvq:324-382: sum over quantizers of codebook_q[indices[..., q]]; index -1 contributes zeros."""
    out = np.zeros((*indices.shape[:-1], embeds[0].shape[-1]), dtype=F32)
    for q, e in enumerate(embeds):
        idx = indices[..., q]
        codes = e[np.where(idx < 0, 0, idx)]
        codes = np.where((idx < 0)[..., None], F32(0), codes)
        out = out + codes
    return out


# --------------------------------------------------------------------------------------------
# tie / near-tie classification used by the parity tests
# --------------------------------------------------------------------------------------------


def top2_gap_f64(x: np.ndarray, embed: np.ndarray, cosine: bool):
    """For each row: (best index in float64 arithmetic, relative gap between the two best scores).

    Rows whose gap is below fp32 rounding noise are 'reference-internal near ties': two correct fp32
    evaluations of the reference formula (MKL vs cuBLAS vs this file) may legitimately disagree on them.
    Scores are squared distances (euclid) or dot products (cosine).
    """
    x64 = x.astype(np.float64)
    e64 = embed.astype(np.float64)
    if cosine:
        s = x64 @ e64.T
        scale = np.ones(x.shape[0])
    else:
        s = -((x64 * x64).sum(-1)[:, None] + (e64 * e64).sum(-1)[None, :] - 2.0 * (x64 @ e64.T))
        scale = np.maximum((x64 * x64).sum(-1), 1e-30)
    part = np.partition(s, -2, axis=-1)
    gap = (part[:, -1] - part[:, -2]) / scale
    return np.argmax(s, axis=-1), gap


# --------------------------------------------------------------------------------------------
# SimVQ (sim_vq.py:99-139): frozen codebook through a linear map; cdist + argmin; two commitment terms
# --------------------------------------------------------------------------------------------
def simvq_forward(x: np.ndarray, frozen: np.ndarray, weight: np.ndarray, input_to_quantize_commit_loss_weight=0.25,
                  commitment_weight=1.0):
    """x (b, n, d) fp32, frozen (K, f) fp32, weight (d, f) of nn.Linear(f, d, bias=False) (sim_vq.py:64).
    Returns (quantized before the gradient estimator — the rotation trick reproduces it numerically, sim_vq.py:126-128 —,
    indices, loss)."""
    x = x.astype(F32)
    codes = (frozen.astype(F32) @ weight.astype(F32).T).astype(F32)           # sim_vq.py:81-83
    flat = x.reshape(-1, x.shape[-1])
    # torch.cdist (sim_vq.py:112) in its euclidean-via-matmul form, like vqp:58-62
    x2 = (flat * flat).sum(-1, dtype=F32)[:, None]
    y2 = (codes * codes).sum(-1, dtype=F32)[None, :]
    d2 = np.maximum((x2 + y2).astype(F32) - (F32(2) * (flat @ codes.T)).astype(F32), F32(0))
    ind = np.argmin(np.sqrt(d2), axis=-1).reshape(x.shape[:-1])               # first minimum, like torch.argmin
    q = codes[ind]
    mse = np.mean((x - q).astype(F32) ** 2, dtype=F32)
    loss = F32(mse + mse * F32(input_to_quantize_commit_loss_weight)) * F32(commitment_weight)   # sim_vq.py:121-124, :139
    return q.astype(F32), ind.astype(np.int64), loss
