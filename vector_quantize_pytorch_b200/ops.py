"""Functional layer over the C ABI: torch tensors in, torch tensors out, everything enqueued on the
current CUDA stream.  Mirrors the arithmetic steps of `Codebook.forward`
(reference vector_quantize_pytorch.py:674-791); the nn.Modules in this package are glue around these.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _C
from ._C import lib, check

# A row is certified by the tensor-core passes when its best score leads all others by more than the band
#   2 * (||x|| * cres + xaux * caux + margin * ||x|| * max||c|| [+ 2^-21 max||c||^2]) + (tag slack, sqrt-collapse width):
# Cauchy-Schwarz on the EXACT norms of what the pass scheme leaves out (csrc/code_operands.cuh, vq_assign.cu) — single fp16
# pass: cres = max_k ||c - fp16 plane||, xaux = norm of the row's flushed elements; bf16 split: cres = max_k ||c - hi - lo||,
# and for fp32 inputs xaux = ||x_lo|| with caux = 2^-8 max||c|| + max||c_lo|| — plus `margin` for the fp32 accumulation in the
# tensor core alone: the TOTAL error of the bf16 split was measured at <= 2^-19.3 ||x|| max||c|| on B200, so 2^-18 for the
# accumulation share keeps > 2.5x.  tests/test_parity_gpu.py::test_score_error_inside_margin asserts the bound for every
# scheme on randn, heavy-tailed, tiny, unit-norm and default-init data.  See DESIGN.md 4.1.
DEFAULT_MARGIN = 2.0 ** -18

_DT = {torch.float32: _C.DTYPE_F32, torch.bfloat16: _C.DTYPE_BF16}

# bench.py instrumentation: when PROFILE_EVENTS is a list, `search` brackets the tcgen05 kernel with CUDA
# events on the launching stream; LAUNCHES counts the kernels this library enqueues.
PROFILE_EVENTS = None
LAUNCHES = 0


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"vqb200 supports float32 and bfloat16 inputs, got {t.dtype}") from None


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vqb200 has no CPU path: tensors must live on a CUDA (B200, sm_100) device")


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def padded_codes(K: int) -> int:
    return lib.vqb_padded_codes(K)


@dataclass
class CodebookOperands:
    """Tensor-core view of one codebook (see vqb_codebook_prepare in include/vqb200.h)."""
    planes: torch.Tensor  # 2-byte (3, Kpad, D): bf16 hi, bf16 lo (bit patterns), fp16(c) (csrc/code_operands.cuh)
    bext: torch.Tensor  # bf16 (Kpad, 16): -bias as three bf16 terms (the operand of the "bias MMA")
    bias: torch.Tensor  # f32 (Kpad,)
    cnorm2: torch.Tensor  # f32 (K,)
    cmax: torch.Tensor  # f32 (4,): max||c||, max||c - fp16 plane||, max||c - bf16 hi - bf16 lo||, max||bf16 lo||
    scratch: torch.Tensor  # f32 (2,)
    K: int
    D: int
    cosine: bool

    @staticmethod
    def allocate(K: int, D: int, cosine: bool, device) -> "CodebookOperands":
        Kpad = padded_codes(K)
        return CodebookOperands(
            planes=torch.empty((3, Kpad, D), dtype=torch.float16, device=device),
            bext=torch.empty((Kpad, 16), dtype=torch.bfloat16, device=device),
            bias=torch.empty((Kpad,), dtype=torch.float32, device=device),
            cnorm2=torch.empty((K,), dtype=torch.float32, device=device),
            cmax=torch.zeros((4,), dtype=torch.float32, device=device),
            scratch=torch.zeros((2,), dtype=torch.float32, device=device),
            K=K, D=D, cosine=cosine)


def prepare_codebook(embed: torch.Tensor, cosine: bool, out: CodebookOperands | None = None) -> CodebookOperands:
    """embed (K, D) fp32 contiguous -> operands for `search`."""
    _require_cuda(embed)
    assert embed.dtype == torch.float32 and embed.dim() == 2 and embed.is_contiguous()
    K, D = embed.shape
    ops = out if out is not None else CodebookOperands.allocate(K, D, cosine, embed.device)
    with torch.cuda.device(embed.device):
        check(lib.vqb_codebook_prepare(_p(embed), K, D, int(cosine), _p(ops.planes), _p(ops.bext), _p(ops.bias), _p(ops.cnorm2),
                                       _p(ops.cmax), _stream()), "vqb_codebook_prepare")
    _count(1)
    return ops


@dataclass
class SearchResult:
    idx: torch.Tensor  # int32 (N,)
    x_eff: torch.Tensor  # (N, D) input as the codebook sees it (l2-normalised for cosine), in x.dtype
    flag_count: torch.Tensor  # int32 (1,) rows re-scored exactly
    flagged: torch.Tensor  # int32 (N, 8): (row, count, cand0, cand1, cand2, pad x 3) — vqb_flag_entry
    best: torch.Tensor | None = None
    rescan_count: torch.Tensor | None = None  # int32 (1,) rows re-scanned whole (entries at the back of `flagged`)


def search(x: torch.Tensor, ops: CodebookOperands, embed: torch.Tensor, *, margin: float | None = None, n_passes: int = 0,
           debug_best: bool = False, fix: bool = True, normalise: bool = True, fused: dict | None = None) -> SearchResult:
    """Nearest code of every row of x (N, D).  Replaces cdist/einsum + argmax (vqp:58-62, :741-747, :130-145).

    fused: optional dict(q_out=, idx64_out=, idx_stride=, loss_sum=, resid_out=, qsum=) of output tensors — the
    gather / commitment-loss / residual tail (see `gather`) then runs INSIDE the search kernel (store warps) and
    the re-score kernels, and no separate gather launch is needed."""
    _require_cuda(x, embed)
    assert x.dim() == 2 and x.is_contiguous()
    N, D = x.shape
    assert D == ops.D
    dt = _dtype_code(x)
    cosine = ops.cosine
    l2 = cosine and normalise  # normalise=False: the caller already applied l2norm (Codebook.forward contract)
    dev = x.device
    margin = DEFAULT_MARGIN if margin is None else margin
    with torch.cuda.device(dev):
        st = _stream()
        if x.dtype == torch.bfloat16:
            if l2:  # l2norm in bf16 (vqp:1159 -> :376); the normalised bf16 rows are the A operand
                x_eff = torch.empty_like(x)
                check(lib.vqb_input_prepare(_p(x), dt, N, D, 1, _p(x_eff), None, 0, st), "vqb_input_prepare")
                _count(1)
            else:
                x_eff = x
            a_planes, n_a = x_eff, 1
        else:
            a_planes = torch.empty((2, N, D), dtype=torch.bfloat16, device=dev)   # bf16 hi / lo split of the fp32 input
            x_eff = torch.empty_like(x) if l2 else x
            check(lib.vqb_input_prepare(_p(x), dt, N, D, int(l2), _p(x_eff) if l2 else None, _p(a_planes), 2, st),
                  "vqb_input_prepare")
            _count(1)
            n_a = 2
        idx = torch.empty((N,), dtype=torch.int32, device=dev)
        flagged = torch.empty((N, 8), dtype=torch.int32, device=dev)
        count = torch.zeros((2,), dtype=torch.int32, device=dev)   # [rows with 2 / 3 candidates, rows re-scanned whole]
        best = torch.empty((N,), dtype=torch.float32, device=dev) if debug_best else None
        fo = None
        if fused is not None:
            fo = _C.FusedOutputs(x_eff=_p(x_eff), embed=_p(embed), q_out=_p(fused.get("q_out")),
                                 idx64_out=_p(fused.get("idx64_out")), idx_stride=int(fused.get("idx_stride", 1)),
                                 loss_sum=_p(fused.get("loss_sum")), x_raw=_p(x) if x_eff is not x else None,
                                 resid_out=_p(fused.get("resid_out")), qsum=_p(fused.get("qsum")), stats_cnt=None, stats_sum=None, dtype=dt)
        fo_ref = ctypes.byref(fo) if fo is not None else None
        prof = PROFILE_EVENTS
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib.vqb_assign_ex(_p(a_planes), n_a, N, D, _p(ops.planes), _p(ops.bext), _p(ops.cmax), ops.K, float(margin),
                                int(n_passes), _p(idx), _p(flagged), _p(count), _p(best), fo_ref, int(cosine),
                                _p(ops.cnorm2), st), "vqb_assign")
        if prof is not None:
            ev1.record()
            prof.append((ev0, ev1))
        _count(1 + 3 * int(fix))
        if fix:
            check(lib.vqb_fix_flagged(_p(x_eff), dt, N, D, _p(embed), _p(ops.cnorm2), ops.K, int(cosine), _p(flagged),
                                      _p(count), _p(idx), fo_ref, st), "vqb_fix_flagged")
    return SearchResult(idx, x_eff, count[:1], flagged, best, count[1:])


# 0: EMA statistics accumulated inside the search kernel (vector RED into L2); 1: separate sort + segmented sums
STATS_MODE = int(__import__("os").environ.get("VQB_STATS_MODE", "1"))
_WS_CACHE: dict = {}


def _workspace(key, nbytes: int, device) -> torch.Tensor:
    """Scratch buffers are reused across calls (same stream => ordered).  Keyed by owner / device only and grown on
    demand: a masked batch that compacts to a different N every call must not pin a new buffer per N."""
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes or buf.device != device:
        buf = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


def vq_forward_args(x: torch.Tensor, ops: CodebookOperands, state: tuple, *, update: int, do_normalise: bool, decay: float,
                    eps: float, q_out=None, idx64_out=None, idx_stride: int = 1, loss_out=None, loss_weight: float = 1.0,
                    resid_out=None, qsum=None, stats=None, margin: float | None = None, already_normalised: bool = False,
                    ws_key=None, stats_accumulate: bool = False, peer=None, peer_ptrs=None, peer_slice_offset: int = 0,
                    a_planes_in=None, planes_out=None, row_mask=None, n_live=None):
    """The argument block of one vqb_vq_forward call (also one VQB_RVQ_STAGE op of vqb_rvq_forward).
    row_mask (N,) uint8 / n_live (1,) int64 on the device: a masked batch (vqp:1116-1119) — padding rows (0) keep the values
    q_out / idx64_out were pre-filled with and stay out of the loss and the statistics (include/vqb200.h).
    Returns (args, idx32, stats, n_launches)."""
    if row_mask is not None:
        assert row_mask.dtype == torch.uint8 and row_mask.is_contiguous() and row_mask.numel() == x.shape[0] and row_mask.is_cuda
        assert n_live is None or (n_live.dtype == torch.int64 and n_live.numel() == 1 and n_live.is_cuda)
    _require_cuda(x, state[2])
    assert x.dim() == 2 and x.is_contiguous()
    N, D = x.shape
    K = ops.K
    dt = _dtype_code(x)
    dev = x.device
    cs, ea, emb = state
    # internal scratch lives in reusable buffers: stable pointers keep the CUDA-graph cache of vqb_vq_forward hot
    idx32 = _workspace(("idx32", ws_key, dev.index), 4 * N, dev)[:4 * N].view(torch.int32)
    if update and stats is None:
        stats = torch.empty((stats_floats(K, D),), dtype=torch.float32, device=dev)
    nbytes = lib.vqb_vq_forward_workspace(N, D, K, dt, int(ops.cosine), int(update))
    ws = _workspace(("fwd", ws_key, dev.index), nbytes, dev)
    a = _C.VQForwardArgs(
        x=_p(x), dtype=dt, metric=int(ops.cosine), N=N, D=D, K=K, already_normalised=int(already_normalised),
        cluster_size=_p(cs), embed_avg=_p(ea), embed=_p(emb), planes=_p(ops.planes), bext=_p(ops.bext), bias=_p(ops.bias),
        cnorm2=_p(ops.cnorm2), cmax=_p(ops.cmax), scratch=_p(ops.scratch), q_out=_p(q_out), idx64_out=_p(idx64_out),
        idx_stride=int(idx_stride), loss_out=_p(loss_out), loss_weight=float(loss_weight), resid_out=_p(resid_out),
        qsum=_p(qsum), idx32=_p(idx32), update=int(update), stats_mode=STATS_MODE, stats_accumulate=int(stats_accumulate), do_normalise=int(do_normalise), decay=float(decay),
        eps=float(eps), stats=_p(stats), margin_rel=float(DEFAULT_MARGIN if margin is None else margin),
        workspace=_p(ws), workspace_bytes=nbytes, ev_search_begin=None, ev_search_end=None,
        a_planes_in=_p(a_planes_in), planes_out=_p(planes_out), row_mask=_p(row_mask), n_live=_p(n_live))
    if update == 3:  # multi-GPU: statistics -> peer barrier -> EMA kernels summing every rank's statistics (vq_peer.cu)
        a.peer_stats = ctypes.cast(peer_ptrs, ctypes.c_void_p)
        a.peer_flags = ctypes.cast(peer.flag_ptrs, ctypes.c_void_p)
        a.peer_epoch = peer.epoch.data_ptr()
        a.peer_rank, a.peer_world, a.peer_slice_offset = peer.rank, peer.world, int(peer_slice_offset)
    n_launch = (4 + (1 if dt == _C.DTYPE_F32 or ops.cosine else 0) + (1 if loss_out is not None else 0) + (5 if update else 0)
                + (2 if update == 2 else 0) + (3 if update == 3 else 0))
    return a, idx32, stats, n_launch


def vq_forward(x: torch.Tensor, ops: CodebookOperands, state: tuple, **kw) -> tuple[torch.Tensor, torch.Tensor | None]:
    """ONE C call for the arithmetic of VectorQuantize.forward / one ResidualVQ stage (vqb_vq_forward).

    state = (cluster_size (K,), embed_avg (K, D), embed (K, D)).  update: 0 none, 1 statistics only (returned
    packed; the caller all-reduces and calls `ema_apply`), 2 statistics + EMA apply.  Returns (idx32, stats)."""
    a, idx32, stats, n_launch = vq_forward_args(x, ops, state, **kw)
    with torch.cuda.device(x.device):
        prof = PROFILE_EVENTS
        if prof is not None:  # bench instrumentation: CUDA events around the search kernel, recorded from C
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); ev1.record()  # materialise the handles
            a.ev_search_begin, a.ev_search_end = ev0.cuda_event, ev1.cuda_event
            prof.append((ev0, ev1))
        check(lib.vqb_vq_forward(ctypes.byref(a), _stream()), "vqb_vq_forward")
    _count(n_launch)
    return idx32, stats


class RvqProgram:
    """The op list of one vqb_rvq_forward call: a whole ResidualVQ / GroupedResidualVQ forward — stages, running sum, deferred
    EMA updates — enqueued by ONE FFI call and replayed from one CUDA graph.  Ops of one lane run in order; lanes (the groups
    of GroupedResidualVQ) run on parallel streams."""
    MAX_OPS = 62

    def __init__(self, device):
        self.device = device
        self.ops = []
        self.keep = []       # tensors the ops point into
        self.launches = 0

    def stage(self, lane, x, cb_ops, state, **kw):
        a, idx32, stats, n = vq_forward_args(x, cb_ops, state, **kw)
        op = _C.RvqOp(kind=_C.RVQ_STAGE, lane=lane)
        op.stage = a
        self.ops.append(op)
        self.keep.append((x, cb_ops, state, kw, idx32, stats))
        self.launches += n
        return idx32, stats

    def ema(self, lane, cluster_size, embed_avg, embed, stats, cb_ops, *, decay, eps, do_lerp, do_normalise, n_lerp=1, slice_stride=0):
        K, D = embed.shape
        op = _C.RvqOp(kind=_C.RVQ_EMA, lane=lane)
        op.ema = _C.RvqEmaArgs(cluster_size=_p(cluster_size), embed_avg=_p(embed_avg), embed=_p(embed), stats=_p(stats), K=K, D=D,
                               decay=float(decay), eps=float(eps), metric=int(cb_ops.cosine), do_lerp=int(do_lerp),
                               do_normalise=int(do_normalise), planes=_p(cb_ops.planes), bext=_p(cb_ops.bext), bias=_p(cb_ops.bias),
                               cnorm2=_p(cb_ops.cnorm2), cmax=_p(cb_ops.cmax), scratch=_p(cb_ops.scratch),
                               n_lerp=int(n_lerp), slice_stride=int(slice_stride))
        self.ops.append(op)
        self.keep.append((cluster_size, embed_avg, embed, stats, cb_ops))
        self.launches += 2

    def barrier(self, lane, peer):
        op = _C.RvqOp(kind=_C.RVQ_BARRIER, lane=lane)
        op.bar = _C.RvqBarArgs(flags=ctypes.cast(peer.flag_ptrs, ctypes.c_void_p), epoch=peer.epoch.data_ptr(), rank=peer.rank,
                               world=peer.world)
        self.ops.append(op)
        self.keep.append(peer)
        self.launches += 1

    def ema_peers(self, lane, cluster_size, embed_avg, embed, peer, peer_ptrs, slice_offset, cb_ops, *, decay, eps, do_normalise,
                  n_lerp=1, slice_stride=0):
        K, D = embed.shape
        op = _C.RvqOp(kind=_C.RVQ_EMA_PEERS, lane=lane)
        op.emap = _C.RvqEmaPeersArgs(cluster_size=_p(cluster_size), embed_avg=_p(embed_avg), embed=_p(embed),
                                     peer_stats=ctypes.cast(peer_ptrs, ctypes.c_void_p), slice_offset=int(slice_offset),
                                     world=peer.world, K=K, D=D, decay=float(decay), eps=float(eps), metric=int(cb_ops.cosine),
                                     do_normalise=int(do_normalise), planes=_p(cb_ops.planes), bext=_p(cb_ops.bext),
                                     bias=_p(cb_ops.bias), cnorm2=_p(cb_ops.cnorm2), cmax=_p(cb_ops.cmax), scratch=_p(cb_ops.scratch),
                                     n_lerp=int(n_lerp), slice_stride=int(slice_stride))
        self.ops.append(op)
        self.keep.append((cluster_size, embed_avg, embed, peer, peer_ptrs, cb_ops))
        self.launches += 2

    def accumulate(self, lane, embeds, indices, out):
        N, Q = indices.shape
        if embeds.dim() == 2:
            K, D = embeds.shape
            stride = 0
        else:
            _, K, D = embeds.shape
            stride = K * D
        op = _C.RvqOp(kind=_C.RVQ_ACCUMULATE, lane=lane)
        op.acc = _C.RvqAccArgs(embeds=_p(embeds), embed_stride=stride, Q=Q, K=K, D=D, idx=_p(indices), N=N, out=_p(out),
                               dtype=_DT[out.dtype])
        self.ops.append(op)
        self.keep.append((embeds, indices, out))
        self.launches += 1

    def freeze(self):
        """Materialise the op array once; afterwards only `arr` is patched (cached programs: residual_vq.py)."""
        n = len(self.ops)
        assert 0 < n <= self.MAX_OPS
        self.arr = (_C.RvqOp * n)(*self.ops)
        self.n = n
        self.ops = None
        self.keep = None     # the owner of a cached program keeps its persistent tensors alive itself (and re-binds the rest)
        return self

    def run(self):
        if getattr(self, "arr", None) is None:
            self.freeze()
        with torch.cuda.device(self.device):
            check(lib.vqb_rvq_forward(ctypes.cast(self.arr, ctypes.c_void_p), self.n, _stream()), "vqb_rvq_forward")
        _count(self.launches)


def gather(x_eff: torch.Tensor, embed: torch.Tensor, idx: torch.Tensor, *, q_out: torch.Tensor | None = None,
           idx64_out: torch.Tensor | None = None, idx_stride: int = 1, loss_sum: torch.Tensor | None = None,
           x_raw: torch.Tensor | None = None, resid_out: torch.Tensor | None = None,
           qsum: torch.Tensor | None = None) -> None:
    """quantize = embed[idx].type(x.dtype) (vqp:766/:779-781, :1178) fused with the mse partial sum (vqp:1327)
    and, for ResidualVQ, residual -= q ; quantized_out += q (rvq:524-525)."""
    _require_cuda(x_eff, embed, idx)
    N, D = x_eff.shape
    with torch.cuda.device(x_eff.device):
        check(lib.vqb_gather(_p(x_eff), _dtype_code(x_eff), N, D, _p(embed), _p(idx), _p(q_out), _p(idx64_out),
                             int(idx_stride), _p(loss_sum), _p(x_raw), _p(resid_out), _p(qsum), _stream()), "vqb_gather")
    _count(1)


def loss_finalize(loss_sum: torch.Tensor, numel: int, dtype: torch.dtype, weight: float, out: torch.Tensor) -> None:
    with torch.cuda.device(loss_sum.device):
        check(lib.vqb_loss_finalize(_p(loss_sum), int(numel), _DT[dtype], float(weight), _p(out), _stream()),
              "vqb_loss_finalize")
    _count(1)


def stats_floats(K: int, D: int) -> int:
    return lib.vqb_stats_floats(K, D)


def stats_offset(K: int) -> int:
    return lib.vqb_stats_offset(K)


def ema_stats(x_eff: torch.Tensor, idx: torch.Tensor, K: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Packed [cluster_size | embed_sum] of this batch (vqp:602, :605) — ready for ONE all-reduce."""
    _require_cuda(x_eff, idx)
    N, D = x_eff.shape
    dev = x_eff.device
    stats = out if out is not None else torch.empty((stats_floats(K, D),), dtype=torch.float32, device=dev)
    ws_bytes = lib.vqb_ema_stats_workspace(N, K)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.vqb_ema_stats(_p(x_eff), _dtype_code(x_eff), N, D, _p(idx), K, _p(stats), _p(ws), ws_bytes, _stream()),
              "vqb_ema_stats")
    _count(4)
    return stats


def ema_apply(cluster_size: torch.Tensor, embed_avg: torch.Tensor, embed: torch.Tensor, stats: torch.Tensor | None,
              ops: CodebookOperands, *, decay: float, eps: float, do_lerp: bool, do_normalise: bool,
              code_weight: torch.Tensor | None = None) -> None:
    """lerp_ of cluster_size / embed_avg (vqp:616-617) and update_ema (vqp:576-584); refreshes `ops`.
    code_weight (K,) fp32: the reference's per-code `ema_update_weight` (vqp:86-97)."""
    _require_cuda(cluster_size, embed_avg, embed, code_weight)
    K, D = embed.shape
    if code_weight is not None:
        assert code_weight.dtype == torch.float32 and code_weight.is_contiguous() and code_weight.numel() == K
    with torch.cuda.device(embed.device):
        check(lib.vqb_ema_apply_weighted(_p(cluster_size), _p(embed_avg), _p(embed), _p(stats), K, D, float(decay), float(eps),
                                         int(ops.cosine), int(do_lerp), int(do_normalise), _p(code_weight), _p(ops.planes),
                                         _p(ops.bext), _p(ops.bias), _p(ops.cnorm2), _p(ops.cmax), _p(ops.scratch), _stream()),
              "vqb_ema_apply")
    _count(2)


def decode(embeds: torch.Tensor, indices: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """sum_q embeds[q][indices[..., q]] with -1 -> zeros (rvq:324-382).  embeds (Q, K, D) fp32 or (K, D)."""
    _require_cuda(embeds, indices)
    if embeds.dim() == 2:
        embeds = embeds.unsqueeze(0)
    Q, K, D = embeds.shape
    assert indices.shape[-1] == Q and indices.dtype == torch.int64
    embeds = embeds.contiguous()
    flat = indices.reshape(-1, Q).contiguous()
    N = flat.shape[0]
    out = torch.empty((N, D), dtype=out_dtype, device=embeds.device)
    with torch.cuda.device(embeds.device):
        check(lib.vqb_decode(_p(embeds), K * D, Q, K, D, _p(flat), N, _p(out), _DT[out_dtype], _stream()), "vqb_decode")
    _count(1)
    return out.reshape(*indices.shape[:-1], D)


def rvq_accumulate(embeds: torch.Tensor, indices: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """ResidualVQ's `quantized_out` from the stage indices (N, Q): the rounded running sum of rvq:525 in ONE pass.
    embeds (Q, K, D) fp32 — the codebooks the stages SEARCHED (pre-update), or (K, D) for a shared codebook."""
    _require_cuda(embeds, indices)
    assert indices.dim() == 2 and indices.dtype == torch.int64 and indices.is_contiguous()
    N, Q = indices.shape
    if embeds.dim() == 2:
        K, D = embeds.shape
        stride = 0
    else:
        assert embeds.shape[0] == Q
        _, K, D = embeds.shape
        stride = K * D
    embeds = embeds.contiguous()
    out = torch.empty((N, D), dtype=out_dtype, device=embeds.device)
    with torch.cuda.device(embeds.device):
        check(lib.vqb_rvq_accumulate(_p(embeds), stride, Q, K, D, _p(indices), N, _p(out), _DT[out_dtype], _stream()),
              "vqb_rvq_accumulate")
    _count(1)
    return out


def peer_barrier(peer) -> None:
    """Cross-GPU barrier kernel on the current stream (dist.PeerReducer; csrc/vq_peer.cu)."""
    with torch.cuda.device(peer.device):
        check(lib.vqb_peer_barrier(ctypes.cast(peer.flag_ptrs, ctypes.c_void_p), peer.rank, peer.world, peer.epoch.data_ptr(),
                                   _stream()), "vqb_peer_barrier")
    _count(1)


def ema_apply_peers(cluster_size: torch.Tensor, embed_avg: torch.Tensor, embed: torch.Tensor, peer, peer_ptrs, slice_offset: int,
                    ops: CodebookOperands, *, decay: float, eps: float, do_normalise: bool,
                    code_weight: torch.Tensor | None = None) -> None:
    """`ema_apply` with the statistics summed over every rank's symmetric buffer inside the kernels (after `peer_barrier`)."""
    _require_cuda(cluster_size, embed_avg, embed, code_weight)
    K, D = embed.shape
    with torch.cuda.device(embed.device):
        check(lib.vqb_ema_apply_peers(_p(cluster_size), _p(embed_avg), _p(embed), ctypes.cast(peer_ptrs, ctypes.c_void_p), peer.world,
                                      int(slice_offset), K, D, float(decay), float(eps), int(ops.cosine), int(do_normalise),
                                      _p(code_weight), _p(ops.planes), _p(ops.bext), _p(ops.bias), _p(ops.cnorm2), _p(ops.cmax),
                                      _p(ops.scratch), _stream()), "vqb_ema_apply_peers")
    _count(2)


def rotate(src: torch.Tensor, tgt: torch.Tensor, grad_out: torch.Tensor | None = None) -> torch.Tensor:
    """Rotation-trick estimator (vqp:287-318): forward value (grad_out None) or the gradient w.r.t. src."""
    _require_cuda(src, tgt, grad_out)
    shape = src.shape
    s2 = src.reshape(-1, shape[-1]).contiguous()
    t2 = tgt.reshape(-1, shape[-1]).contiguous()
    g2 = grad_out.reshape(-1, shape[-1]).contiguous() if grad_out is not None else None
    assert s2.dtype == t2.dtype and (g2 is None or g2.dtype == s2.dtype)
    out = torch.empty_like(s2)
    with torch.cuda.device(s2.device):
        check(lib.vqb_rotate(_p(s2), _p(t2), _p(g2), s2.shape[0], s2.shape[1], _dtype_code(s2), _p(out), _stream()), "vqb_rotate")
    _count(1)
    return out.reshape(shape)
