"""Second CPU restatement of the hot path using torch CPU ops — TEST / BASELINE INFRASTRUCTURE ONLY.

Same status as oracle/vq_oracle.py (never imported by the product).  It restates
`Codebook.forward` + the tail of `VectorQuantize.forward` (vector_quantize_pytorch.py:674-791, :1178,
:1327) with the very ATen operators the reference calls (einsum -> MKL sgemm, argmax, one_hot, lerp_), so
that (a) `bench.py --impl reference` / `cpu_baseline` time the reference's real CPU cost — the (N x K)
fp32 distance matrix, the int64 -> fp32 one-hot and three dense GEMMs — and (b) the numpy oracle has an
independent cross-check.  Pinned by tests/test_oracle_golden.py against the same golden fixtures.
"""
import torch
import torch.nn.functional as F


def l2norm(t):
    return F.normalize(t, p=2, dim=-1, eps=1e-6)  # vqp:37-38


def cdist(x, y, eps=1e-8):  # vqp:58-62
    x2 = (x ** 2).sum(-1)
    y2 = (y ** 2).sum(-1)
    xy = torch.einsum("b i d, b j d -> b i j", x, y) * -2
    return (x2[:, :, None] + y2[:, None, :] + xy).clamp(min=eps).sqrt()


class State:
    def __init__(self, embed):
        self.embed = embed.clone().float()[None]          # (1, K, D)   vqp:423
        self.embed_avg = self.embed.clone()               # vqp:417
        self.cluster_size = torch.ones(1, embed.shape[0]) # vqp:416


@torch.no_grad()
def vq_forward(x, state, *, cosine=False, training=True, decay=0.8, eps=1e-5, commitment_weight=1.0,
               manual_ema_update=False):
    """x (..., D) in its own dtype (fp32 / bf16).  Returns (quantize, indices int64, loss fp32)."""
    dtype = x.dtype
    if cosine:
        x = l2norm(x)                                      # vqp:1159 (input dtype)
    flatten = x.float().reshape(1, -1, x.shape[-1])        # vqp:692-698
    embed = state.embed
    K = embed.shape[1]
    dist = torch.einsum("h n d, h c d -> h n c", flatten, embed) if cosine else -cdist(flatten, embed)  # vqp:741/:743
    ind = dist.argmax(dim=-1)                              # vqp:140
    onehot = F.one_hot(ind, K).type(torch.float32)         # vqp:142
    if training:
        quantize = torch.einsum("h n c, h c d -> h n d", onehot, embed)      # vqp:766
    else:
        quantize = embed[0][ind[0]][None]                  # vqp:779-781
    if training:                                           # vqp:586-617
        cluster_size = onehot.sum(dim=1)
        embed_sum = torch.einsum("h n d, h n c -> h c d", flatten, onehot).contiguous()
        state.cluster_size.lerp_(cluster_size, 1. - decay)
        state.embed_avg.lerp_(embed_sum, 1. - decay)
        if not manual_ema_update:                          # vqp:576-584
            cs = (state.cluster_size + eps) / (state.cluster_size.sum(-1, keepdim=True) + K * eps) * state.cluster_size.sum(-1, keepdim=True)
            e = state.embed_avg / cs[..., None]
            state.embed.copy_(l2norm(e) if cosine else e)
    quantize = quantize.reshape(x.shape).type(dtype)       # vqp:1178
    loss = torch.tensor(0.)
    if training and commitment_weight > 0:
        loss = loss + F.mse_loss(quantize, x) * commitment_weight   # vqp:1327-1329
    return quantize, ind.reshape(x.shape[:-1]), loss
