"""Parity at the BASELINE shapes: fixtures generated from the UNMODIFIED reference (oracle/gen_golden_big.py) at
cfg2 (D=256, K=1024, 32768 rows; bf16 / fp32; warm and default-init codebooks), cfg3 (ResidualVQ Q=8 shared, 8192 rows),
cfg4 (cosine D=512 K=16384, 4096 rows; bf16 / fp32) and cfg5 (GroupedResidualVQ G=2 Q=8, 8192 rows).

The inputs are regenerated from the seeded recipe (oracle/big_cases.py) and checked against the stored checksums.
CPU leg: the numpy oracle replays every fixture.  GPU leg (`-m gpu`): the CUDA modules replay them.
Bar: indices identical to the reference's.  A row may differ only if the reference's own decision is inside fp32
rounding noise (float64 top-2 gap of the reference formula below `tie_tol`); such rows are listed and bounded (<= 1e-4
of the rows; none at all expected for the warm VectorQuantize fixtures of step 0).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import big_cases as B
from oracle import vq_oracle as O

BIG = sorted(B.CASES)


class BigGolden:
    def __init__(self, name):
        self.name = name
        self.case = B.CASES[name]
        self.z = np.load(os.path.join(B.GOLDEN, name + ".npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())

    def inputs(self):
        inits = B.initial_states(self.case)
        xs = B.step_inputs(self.case)
        assert [B.digest(e) for e, _, _ in inits] == self.meta["init_digest"], "initial codebooks drifted from the fixture's"
        assert [B.digest(x) for x in xs] == self.meta["x_digest"], "regenerated inputs drifted from the fixture's"
        return inits, xs

    @property
    def cfg(self):
        kw = self.case["kw"]
        return O.VQConfig(dim=B.codebook_dim(self.case), codebook_size=kw["codebook_size"],
                          use_cosine_sim=kw.get("use_cosine_sim", False))


def _tie_rows_vq(x, embed, cosine, dtype, tol):
    xx = O.cast_like(x, dtype)
    if cosine:
        xx = O.l2norm(xx, dtype)
    _, gap = O.top2_gap_f64(xx, embed, cosine)
    return gap < tol


def check_step(g, s, q, ind, loss, states_flat, *, pre_embed=None, x=None, who=""):
    """Compare one step's outputs / post-state with the fixture.  Returns the number of excused near-tie rows."""
    z, case = g.z, g.case
    dtype = case["dtype"]
    ref_ind = z[f"s{s}_indices"].astype(np.int64)
    ind = np.asarray(ind).reshape(ref_ind.shape)
    mism = ind != ref_ind
    n_rows = int(np.prod(case["shape"][:-1]))
    excused = 0
    if mism.any():
        if case["kind"] == "vq":
            rows = np.nonzero(mism.reshape(-1))[0]
            xr = x.reshape(-1, x.shape[-1])[rows]
            # fp32 evaluation noise of the reference formula: ~1e-6 relative (warm) — the default-init codebook sits on
            # sqrt-collapsed fp32 ties where the reference's own sgemm rounding decides: wider class
            tol = 2e-6 if case["init"] != "default" else 2e-5
            tie = _tie_rows_vq(xr, pre_embed, case["kw"].get("use_cosine_sim", False), dtype, tol)
            assert tie.all(), f"{who}{g.name} step {s}: {int((~tie).sum())} NON-tie index mismatches (rows {rows[~tie][:8]})"
            excused = int(mism.sum())
        else:  # residual stacks: one flipped near tie changes every later stage of that row — count rows
            excused = int(mism.reshape(n_rows, -1).any(-1).sum())
        print(f"{who}{g.name} step {s}: {excused} excused near-tie rows of {n_rows}")
        assert excused <= max(1, int(1e-4 * n_rows)) or (case["init"] == "default" and s == 0 and excused <= 2e-3 * n_rows), \
            f"{who}{g.name} step {s}: {excused} mismatching rows"
    vt = 1e-5 if dtype == "fp32" else 8e-3
    ok_rows = ~mism.reshape(n_rows, -1).any(-1)
    qf = np.asarray(q, dtype=np.float32).reshape(n_rows, -1)
    sel = ok_rows[:64]
    np.testing.assert_allclose(qf[:64][sel], z[f"s{s}_q_rows"][sel], rtol=vt, atol=vt)
    if not mism.any():
        np.testing.assert_allclose(qf.astype(np.float64).sum(0), z[f"s{s}_q_colsum"], rtol=1e-4, atol=2e-2 if dtype == "fp32" else 2.0)
    np.testing.assert_allclose(np.asarray(loss, dtype=np.float32).reshape(-1), z[f"s{s}_loss"].reshape(-1),
                               rtol=2e-5 if dtype == "fp32" else 8e-3, atol=1e-7)
    rows = B.sample_rows(case["kw"]["codebook_size"])
    for i, (embed, embed_avg, cs) in enumerate(states_flat):
        # a flipped near-tie row moves one count between two codes: compare up to that
        tol_cs = 0.2 * excused + 1e-4
        np.testing.assert_allclose(cs, z[f"s{s}_cb{i}_cluster_size"], rtol=1e-5, atol=tol_cs)
        if excused == 0:
            # step >= 1 of a bf16 case starts from OUR step-0 codebook (equal to the reference's to ~5e-7): an element that
            # sits on a bf16 rounding boundary of `quantize = embed.type(bf16)` (vqp:1178) then moves by one bf16 ulp and
            # with it one term of the residual statistics (observed: 4 of 16128 sampled elements, 1e-4 absolute)
            at = 2e-5 if (s == 0 or dtype == "fp32") else 3e-4
            np.testing.assert_allclose(embed[rows], z[f"s{s}_cb{i}_embed_rows"], rtol=2e-5, atol=at)
            np.testing.assert_allclose(embed_avg[rows], z[f"s{s}_cb{i}_embed_avg_rows"], rtol=2e-5, atol=10 * at)
            np.testing.assert_allclose(embed.astype(np.float64).sum(0), z[f"s{s}_cb{i}_embed_colsum"], rtol=1e-4, atol=1e-3 if at < 1e-4 else 2e-2)
    return excused


# ---------------------------------------------------------------------------------------------------------- CPU: oracle
def _oracle_states(g, inits):
    mk = lambda t: O.CodebookState(t[0].numpy().copy(), t[1].numpy().copy(), t[2].numpy().copy())
    flat = [mk(t) for t in inits]
    kw = g.case["kw"]
    if g.case["kind"] == "vq":
        return flat[0], flat
    Q = kw["num_quantizers"]
    if g.case["kind"] == "rvq":
        return ([flat[0]] * Q if kw.get("shared_codebook") else flat), flat
    G = kw["groups"]
    if kw.get("shared_codebook"):
        return [[flat[i]] * Q for i in range(G)], flat
    return [flat[i * Q:(i + 1) * Q] for i in range(G)], flat


@pytest.mark.parametrize("name", BIG)
def test_oracle_replays_big_fixture(name):
    g = BigGolden(name)
    inits, xs = g.inputs()
    states, flat = _oracle_states(g, inits)
    kw, dtype = g.case["kw"], g.case["dtype"]
    for s, x in enumerate(xs):
        xn = x.float().numpy()
        pre = flat[0].embed.copy()
        if g.case["kind"] == "vq":
            q, ind, loss, _ = O.vq_forward(xn, dtype, states, g.cfg)
        elif g.case["kind"] == "rvq":
            q, ind, loss, _ = O.rvq_forward(xn, dtype, states, g.cfg, shared_codebook=kw.get("shared_codebook", False))
        else:
            q, ind, loss, _ = O.grouped_rvq_forward(xn, dtype, states, g.cfg, shared_codebook=kw.get("shared_codebook", False))
        if check_step(g, s, q, ind, loss, [(st.embed, st.embed_avg, st.cluster_size) for st in flat], pre_embed=pre, x=xn,
                      who="oracle: "):
            break  # a flipped near tie changes the codebook: later steps no longer start from the reference's state


# ---------------------------------------------------------------------------------------------------------- GPU: product
def _ours_codebooks(module):
    import vector_quantize_pytorch_b200 as m
    seen, out = set(), []
    for sub in module.modules():
        if isinstance(sub, m.Codebook) and id(sub) not in seen:
            seen.add(id(sub))
            out.append(sub)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", BIG)
def test_cuda_modules_replay_big_fixture(name):
    import vector_quantize_pytorch_b200 as m
    g = BigGolden(name)
    inits, xs = g.inputs()
    kw = g.case["kw"]
    cls = {"vq": m.VectorQuantize, "rvq": m.ResidualVQ, "grvq": m.GroupedResidualVQ}[g.case["kind"]]
    module = cls(**kw).to("cuda:0")
    books = _ours_codebooks(module)
    assert len(books) == len(inits)
    with torch.no_grad():
        for cb, (e, ea, cs) in zip(books, inits):
            cb.embed.copy_(e[None]); cb.embed_avg.copy_(ea[None]); cb.cluster_size.copy_(cs[None])
    module.train()
    for s, x in enumerate(xs):
        pre = books[0].embed[0].cpu().numpy().copy()
        q, ind, loss = module(x.to("cuda:0"))[:3]
        torch.cuda.synchronize()
        assert q.dtype == x.dtype and ind.dtype == torch.int64
        flat = [(cb.embed[0].cpu().numpy(), cb.embed_avg[0].cpu().numpy(), cb.cluster_size[0].cpu().numpy()) for cb in books]
        if check_step(g, s, q.float().cpu().numpy(), ind.cpu().numpy(), loss.detach().float().cpu().numpy(), flat,
                      pre_embed=pre, x=x.float().numpy(), who="cuda: "):
            break  # a flipped near tie changes the codebook: later steps no longer start from the reference's state
