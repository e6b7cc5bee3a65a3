"""CPU model of the search kernel's epilogue arithmetic (vq_assign.cu: RowState::piece, the slice merge).

The CUDA epilogue keeps, per row, a branch-free top-3 of *tagged* scores (4 low mantissa bits = 15 - column inside the
16-column piece), skips pieces whose exact maximum is below the running threshold, recovers every compare-exchange
minimum with the integer identity min = a + b - max, and merges the two column slices of a row.  This file restates that
arithmetic bit for bit in numpy and checks the CERTIFICATE the parity argument rests on against brute
force on adversarial score matrices:

  * every code NOT reported as a candidate scores at least W - 2*(tag slack) below the exact row maximum, so the exact
    arg-max is always among the reported candidates (1 candidate: certified; 2: exact pair re-score; >= 3: whole-row rescan);
  * reported indices are valid and distinct; the exact winning score (`bexact`, feeds the loss) is the true maximum.

It is a model of the algorithm (test infrastructure), not of the GPU: the CUDA code itself is checked on the device by
tests/test_parity_gpu.py.
"""
import numpy as np
import pytest

NEG = np.float32(-3.4e38)


def f2u(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def u2f(x):
    return np.ascontiguousarray(x, dtype=np.uint32).view(np.float32)


class Slice:
    """RowState of one column slice, vectorised over rows (vq_assign.cu: struct RowState)."""

    def __init__(self, W):
        R = W.shape[0]
        self.W = W.astype(np.float32)
        self.t1 = np.full(R, NEG, np.float32)
        self.t2 = np.full(R, NEG, np.float32)
        self.t3 = np.full(R, NEG, np.float32)
        self.thr = np.full(R, NEG, np.float32)
        self.bexact = np.full(R, NEG, np.float32)
        self.j1 = np.zeros(R, np.int64)
        self.j2 = np.zeros(R, np.int64)

    def piece(self, r, cbase):
        """r: (R, 16) float32 scores of one piece; cbase: first column of the piece."""
        mm = r.max(axis=1)
        self.bexact = np.maximum(self.bexact, mm)
        act = mm > self.thr                       # `if (mm > st.thr) st.piece(...)`
        if not act.any():
            return
        t1, t2, t3 = self.t1.copy(), self.t2.copy(), self.t3.copy()
        o1, o2 = t1.copy(), t2.copy()
        bits = f2u(r)
        with np.errstate(over="ignore"):
            for e in range(16):
                ku = (bits[:, e] & np.uint32(0xFFFFFFF0)) | np.uint32(15 - e)
                n1 = np.maximum(t1, u2f(ku))
                lo1 = (f2u(t1) + ku - f2u(n1)).astype(np.uint32)          # min(t1, k) = t1 + k - max(t1, k), mod 2^32
                n2 = np.maximum(t2, u2f(lo1))
                lo2 = (f2u(t2) + lo1 - f2u(n2)).astype(np.uint32)
                t3 = np.maximum(t3, u2f(lo2))
                t1, t2 = n1, n2
        c1 = t1 != o1
        j2 = np.where(t2 == o2, self.j2, np.where(c1 & (t2 == o1), self.j1, cbase))
        j1 = np.where(c1, cbase, self.j1)
        self.t1 = np.where(act, t1, self.t1)
        self.t2 = np.where(act, t2, self.t2)
        self.t3 = np.where(act, t3, self.t3)
        self.j1 = np.where(act, j1, self.j1)
        self.j2 = np.where(act, j2, self.j2)
        self.thr = np.where(act, t1 - self.W, self.thr).astype(np.float32)


def col(t, j):
    return j + 15 - (f2u(t) & np.uint32(15)).astype(np.int64)


def epilogue(V, W, BN=256):
    """(i0, i1, n, best) per row, as the kernel's merge produces them.  V: (R, Kpad) float32, Kpad % BN == 0."""
    R, Kpad = V.shape
    halves = [Slice(W), Slice(W)]
    for ct in range(Kpad // BN):
        for p in range(BN // 16):
            h = (p % 4) // 2                       # pieces 4q + 2*half + {0, 1} belong to column half `half`
            c0 = ct * BN + p * 16
            halves[h].piece(V[:, c0:c0 + 16], c0)
    a, b = halves
    ia0, ia1, ib0, ib1 = col(a.t1, a.j1), col(a.t2, a.j2), col(b.t1, b.j1), col(b.t2, b.j2)
    best = np.maximum(a.bexact, b.bexact)
    tb = np.maximum(a.t1, b.t1)
    band = (tb - a.W).astype(np.float32)
    n = sum((t > band).astype(np.int64) for t in (a.t1, a.t2, a.t3, b.t1, b.t2, b.t3))
    first = (a.t1 > b.t1) | ((a.t1 == b.t1) & (ia0 < ib0))
    i0 = np.where(first, ia0, ib0)
    i1 = np.where(first, np.where(a.t2 > b.t1, ia1, ib0), np.where(b.t2 > a.t1, ib1, ia0))
    return i0, i1, n, best


def make_scores(kind, R, K, rng):
    V = rng.standard_normal((R, K)).astype(np.float32) * np.float32(16.0) - np.float32(100.0)
    if kind == "mixed_sign":
        V = rng.standard_normal((R, K)).astype(np.float32) * np.float32(3.0)
    elif kind == "near_ties":          # several codes within a few float32 ulps .. 1e-3 of the row maximum
        top = V.max(axis=1, keepdims=True)
        for _ in range(4):
            cols = rng.integers(0, K, size=R)
            eps = (rng.random(R).astype(np.float32) ** 4) * np.float32(2e-3)
            V[np.arange(R), cols] = (top[:, 0] - eps * np.abs(top[:, 0])).astype(np.float32)
    elif kind == "exact_ties":         # exact duplicates of the maximum, in the same piece and in other pieces
        top = V.max(axis=1)
        for _ in range(3):
            cols = rng.integers(0, K, size=R)
            V[np.arange(R), cols] = top
    elif kind == "ascending":          # every element is a new maximum: the most updates possible
        V = np.sort(V, axis=1)
    elif kind == "descending":
        V = -np.sort(-V, axis=1)
    elif kind == "constant":
        V[:] = np.float32(-7.25)
    return V


@pytest.mark.parametrize("kind", ["random", "mixed_sign", "near_ties", "exact_ties", "ascending", "descending", "constant"])
@pytest.mark.parametrize("K,BN", [(1024, 256), (300, 256), (48, 48)])
@pytest.mark.parametrize("w_rel", [0.0, 2.0 ** -16, 2.0 ** -10])
def test_epilogue_certificate(kind, K, BN, w_rel):
    rng = np.random.default_rng(hash((kind, K, int(w_rel * 2 ** 20))) % (2 ** 32))
    R = 384
    V = make_scores(kind, R, K, rng)
    Kpad = -(-K // BN) * BN
    Vp = np.full((R, Kpad), np.float32(-3e38), np.float32)   # padded codes: bias -3e38, never candidates
    Vp[:, :K] = V
    vmax_abs = np.abs(V).max(axis=1)
    slack = (np.float32(2.0 ** -18) * vmax_abs).astype(np.float32)        # 2 * (16 ulp <= 2^-19 |score|)
    W = (np.float32(w_rel) * vmax_abs + slack + np.float32(1e-30)).astype(np.float32)
    i0, i1, n, best = epilogue(Vp, W, BN)

    exact_best = V.max(axis=1)
    assert np.array_equal(best, exact_best), "bexact must be the exact row maximum"
    assert ((i0 >= 0) & (i0 < K)).all(), "winner must be a real code"
    two = n == 2
    assert ((i1[two] >= 0) & (i1[two] < K) & (i1[two] != i0[two])).all(), "second candidate must be a distinct real code"

    # the certificate: a code that is not a reported candidate is at least (W - slack) below the exact maximum
    V64 = V.astype(np.float64)
    lim = exact_best.astype(np.float64) - (W.astype(np.float64) - slack.astype(np.float64))
    reported = np.zeros((R, K), bool)
    reported[np.arange(R), i0] = True
    reported[two, i1[two]] = True
    many = n >= 3                                  # whole-row rescan: nothing to certify
    unreported_high = (V64 > lim[:, None]) & ~reported & ~many[:, None]
    assert not unreported_high.any(), f"{int(unreported_high.any(axis=1).sum())} rows hide a candidate from the re-score"
    # hence the exact arg-max is always among the candidates handed on
    am = V64.argmax(axis=1)
    ok = many | (am == i0) | (two & (am == i1)) | (V64[np.arange(R), i0] == V64[np.arange(R), am])
    assert ok.all()
    # certified rows have a strict, unique exact maximum at i0
    one = n == 1
    others = V64.copy()
    others[np.arange(R), i0] = -np.inf
    assert (others[one].max(axis=1) < V64[np.arange(R), i0][one]).all()


def test_constant_rows_go_to_the_whole_row_rescan():
    V = np.full((8, 256), np.float32(1.5), np.float32)
    W = np.full(8, np.float32(1e-6), np.float32)
    _, _, n, best = epilogue(V, W, 256)
    assert (n >= 3).all() and (best == np.float32(1.5)).all()
