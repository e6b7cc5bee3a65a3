"""CPU model of the certification band's reference-facing terms (csrc/vq_assign.cu, `sc.init(...)`).

The search kernel certifies a row when its best tensor-core score leads every other score by more than
    W = 2 (||x|| cres + ||x_lo|| caux + m ||x|| cmax + 2^-21 cmax^2)          <- what the MMA passes can be off by (GPU test:
                                                                                 test_score_error_inside_margin)
      + 2^-18 (||x|| cmax + cmax^2 / 2) + 2^-22 (||x||^2 + cmax^2)            <- tag slack + the width over which the REFERENCE's
                                                                                 own fp32 formula (vqp:58-62: sqrt of a rounded
                                                                                 d^2, first maximal index) departs from exact
and sends every other row to the exact re-score, which evaluates the reference formula itself.  The first line is a property
of the tensor core and is asserted on the GPU.  The second line is a property of the reference's arithmetic and can be checked
here: whenever the reference's fp32 arg-max differs from the exact (float64) arg-max of x.c - |c|^2/2, the exact score gap
between the two winners must lie inside that slack — then the kernel cannot certify such a row with the non-reference index
(round-1 VERDICT weak #2: with the default-init codebook 4 of 32768 rows escaped the old band).
"""
import numpy as np
import pytest
import torch

from oracle import vq_oracle as O


def reference_slack(x2, cmax):
    """The band's last two terms, float64 evaluation of the fp32 expression in vq_assign.cu (Euclidean metric)."""
    xc = np.sqrt(x2) * cmax
    return 2.0 ** -18 * (xc + 0.5 * cmax * cmax) + 2.0 ** -22 * (x2 + cmax * cmax)


def codebook(kind, K, D):
    if kind == "default":  # vqp:112-115, :385: kaiming_uniform_ on the (1, K, D) tensor -> |c| <= sqrt(6 / (K D))
        e = torch.empty(1, K, D)
        torch.nn.init.kaiming_uniform_(e)
        return e[0].numpy()
    if kind == "warm":
        return torch.randn(K, D).numpy()
    if kind == "tiny_and_huge":  # norms spread over four decades
        return (torch.randn(K, D) * torch.logspace(-2, 2, K)[:, None]).numpy()
    raise ValueError(kind)


def rows(kind, N, D):
    x = torch.randn(N, D)
    if kind == "heavy":  # one dominant coordinate per row
        x[torch.arange(N), torch.randint(0, D, (N,))] *= 300.0
    elif kind == "scaled":
        x = x * torch.logspace(-2, 2, N)[:, None]
    return x.numpy().astype(np.float32)


CASES = [
    # K,    D,   N,     codebook,        rows
    (1024, 256, 16384, "default", "randn"),     # BASELINE config 2, first training step of a default-constructed module
    (1024, 256, 8192, "default", "scaled"),
    (1024, 256, 8192, "default", "heavy"),
    (1024, 256, 8192, "warm", "randn"),
    (1024, 256, 4096, "warm", "scaled"),
    (1024, 128, 8192, "default", "randn"),      # config 5 stage shape
    (4096, 512, 2048, "default", "randn"),      # towards config 4's size (Euclidean variant)
    (333, 64, 8192, "tiny_and_huge", "randn"),
    (96, 64, 8192, "default", "randn"),         # the shape of the small cold-init fixture
]


@pytest.mark.parametrize("K,D,N,cb_kind,row_kind", CASES)
def test_reference_departures_lie_inside_the_band(K, D, N, cb_kind, row_kind):
    torch.manual_seed(K * 7 + D + N)
    e = codebook(cb_kind, K, D)
    x = rows(row_kind, N, D)
    ref = O.argmax_first(O.neg_cdist(x, e))                     # the reference's fp32 formula and tie rule (vqp:58-62, :140)
    x64, e64 = x.astype(np.float64), e.astype(np.float64)
    s = x64 @ e64.T - 0.5 * (e64 * e64).sum(-1)[None]           # the kernel's score, exact
    best = s.argmax(-1)
    ar = np.arange(N)
    gap = s[ar, best] - s[ar, ref]
    x2 = (x64 * x64).sum(-1)
    cmax = float(np.sqrt((e64 * e64).sum(-1).max()))
    slack = reference_slack(x2, cmax)
    differing = ref != best
    # every departure of the reference from the exact arg-max is a near tie well inside the slack (factor 1.5 in hand)
    assert (gap[differing] * 1.5 <= slack[differing]).all(), (
        f"{int(differing.sum())} rows differ, worst gap / slack = {(gap[differing] / slack[differing]).max():.3f}")
    # and these two terms alone do not send a trained codebook's rows to the re-score wholesale
    top2 = np.partition(s, -2, axis=-1)[:, -2:]
    frac = float(((top2[:, 1] - top2[:, 0]) < slack).mean())
    if cb_kind == "warm" and row_kind == "randn":
        assert frac < 1e-3, frac
