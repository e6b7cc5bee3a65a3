"""Minimal stand-in for the `einx` package (TEST INFRASTRUCTURE ONLY).

The reference (lucidrains/vector-quantize-pytorch v1.31.0) imports `einx` at module top
(vector_quantize_pytorch/vector_quantize_pytorch.py:16, residual_vq.py:19-20) but `einx` is not
installed in this image and cannot be fetched (no network).  The hot path never calls it; the
only call sites are the mask / decode / beam helpers.  This shim implements exactly the patterns
those call sites use, with plain torch indexing, so that `oracle/ref_loader.py` can import the
UNMODIFIED reference from /root/reference to generate golden vectors.

It is never imported by the product package.
"""
import torch


def _norm(pattern):
    return " ".join(pattern.replace(",", " , ").split())


def where(pattern, cond, a, b):
    p = _norm(pattern)
    if not torch.is_tensor(cond):
        raise TypeError("einx shim: cond must be a tensor")

    def as_t(v, like):
        return v if torch.is_tensor(v) else torch.as_tensor(v, dtype=like.dtype, device=like.device)

    # vector_quantize_pytorch.py:1384  'b n, b n ... d, b n d -> b n ... d'
    if p == _norm("b n, b n ... d, b n d -> b n ... d"):
        extra = a.ndim - 3
        c = cond.reshape(*cond.shape, *((1,) * (extra + 1)))
        bb = b.reshape(*b.shape[:2], *((1,) * extra), b.shape[-1])
        return torch.where(c, a, bb)
    # vector_quantize_pytorch.py:1391  'b n, b n ..., -> b n ...'
    if p == _norm("b n, b n ..., -> b n ..."):
        extra = a.ndim - 2
        c = cond.reshape(*cond.shape, *((1,) * extra))
        return torch.where(c, a, as_t(b, a))
    # vector_quantize_pytorch.py:1315  '..., ... k, -> ... k'
    if p == _norm("..., ... k, -> ... k"):
        return torch.where(cond.unsqueeze(-1), a, as_t(b, a))
    # residual_vq.py:579  '..., ... l,'
    if p == _norm("..., ... l,"):
        return torch.where(cond.unsqueeze(-1), a, as_t(b, a))
    raise NotImplementedError(f"einx shim: where pattern {pattern!r}")


def add(pattern, a, b):
    p = _norm(pattern)
    # residual_vq.py:515  '... j, ... j k -> ... (j k)'
    if p == _norm("... j, ... j k -> ... (j k)"):
        out = a.unsqueeze(-1) + b
        return out.reshape(*out.shape[:-2], -1)
    raise NotImplementedError(f"einx shim: add pattern {pattern!r}")


def get_at(pattern, table, indices):
    p = _norm(pattern)
    # residual_vq.py:346  'q [c] d, b n q -> q b n d'
    if p == _norm("q [c] d, b n q -> q b n d"):
        q = table.shape[0]
        idx = indices.permute(2, 0, 1)                      # q b n
        ar = torch.arange(q, device=table.device).reshape(q, 1, 1)
        return table[ar, idx]
    # residual_vq.py:362 / sim_vq.py:117  '[c] d, b n -> b n d' ; sim_vq.py:92 '[c] d, b ... -> b ... d'
    if p in (_norm("[c] d, b n -> b n d"), _norm("[c] d, b ... -> b ... d")):
        return table[indices]
    # residual_vq.py:360  'b n [c] d, b n -> b n d'
    if p == _norm("b n [c] d, b n -> b n d"):
        d = table.shape[-1]
        idx = indices[..., None, None].expand(*indices.shape, 1, d)
        return table.gather(-2, idx).squeeze(-2)
    raise NotImplementedError(f"einx shim: get_at pattern {pattern!r}")
