"""The numpy oracle must reproduce every golden fixture generated from the real reference.

Bar: indices identical except on rows the reference itself resolves inside fp32 rounding noise
(float64 top-2 gap < 2e-6 relative; counted and bounded); values within 1e-5 (fp32) / one bf16 ulp.
"""
import numpy as np
import pytest

from golden_util import GOLDEN_DIR, Golden, golden_names, near_tie_rows, cpu_pick_fn, simvq_golden_names, mask_golden_names, layout_golden_names, dropout_golden_names
from oracle import vq_oracle as O


def run_oracle(g, step, states, faithful=False):
    m, cfg = g.meta, g.cfg
    x = g[f"s{step}_x"]
    training = m["steps"][step] == "train"
    if m.get("threshold_ema_dead_code", 0) > 0 or m.get("kmeans_init"):  # replay the reference's RNG stream (gen_golden.py seeds every step)
        import torch
        torch.manual_seed(5000 + step)
    if m["kind"] == "vq":
        q, ind, loss, loss32 = O.vq_forward(x, m["dtype"], states, cfg, training=training, faithful=faithful, pick_fn=cpu_pick_fn)
    elif m["kind"] == "rvq":
        q, ind, loss, loss32 = O.rvq_forward(x, m["dtype"], states, cfg, shared_codebook=m["shared_codebook"],
                                              training=training, faithful=faithful, pick_fn=cpu_pick_fn)
    else:
        q, ind, loss, loss32 = O.grouped_rvq_forward(x, m["dtype"], states, cfg, shared_codebook=m["shared_codebook"],
                                                      training=training, faithful=faithful)
    return q, ind, loss, loss32


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("faithful", [False, True])
def test_oracle_matches_reference(name, faithful):
    g = Golden(name)
    m = g.meta
    states = g.states("s0_pre")
    vtol = 1e-5 if m["dtype"] == "fp32" else 8e-3
    for step in range(len(m["steps"])):
        q, ind, loss, _ = run_oracle(g, step, states, faithful)
        ref_ind = g[f"s{step}_indices"]
        mism = ind != ref_ind
        if m["kind"] == "vq" and mism.any() and m.get("heads", 1) == 1:
            # every disagreement must be a reference-internal near tie
            pre = g.state("s0_pre", 0) if step == 0 else g.state(f"s{step - 1}_post", 0)
            x = O.cast_like(g[f"s{step}_x"], m["dtype"]).reshape(-1, m["dim"])
            if m.get("use_cosine_sim"):
                x = O.l2norm(x, m["dtype"])
            tie = near_tie_rows(x, pre.embed, m.get("use_cosine_sim", False)).reshape(mism.shape)
            assert not (mism & ~tie).any(), f"{name} step {step}: non-tie index mismatch"
        if "coldinit" not in name:
            assert mism.sum() == 0, f"{name} step {step}: {mism.sum()} index mismatches"
        else:
            assert mism.mean() < 0.02
            # make the later comparisons meaningful: continue from the reference's state
            states = g.states(f"s{step}_post")
            continue
        np.testing.assert_allclose(q, g[f"s{step}_quantize"], rtol=vtol, atol=vtol)
        np.testing.assert_allclose(loss, g[f"s{step}_loss"], rtol=1e-5 if m["dtype"] == "fp32" else 8e-3, atol=1e-7)
        for i, st in enumerate(g.flat_states(states)):
            ref = g.state(f"s{step}_post", i)
            np.testing.assert_allclose(st.cluster_size, ref.cluster_size, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(st.embed_avg, ref.embed_avg, rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(st.embed, ref.embed, rtol=1e-5, atol=1e-5)


def test_bf16_round_matches_torch():
    import torch
    a = np.random.default_rng(0).standard_normal(100000).astype(np.float32) * 37.0
    ref = torch.from_numpy(a).bfloat16().float().numpy()
    assert np.array_equal(O.bf16_round(a), ref)


def test_decode_invariant():
    """tests/test_readme.py:74-103 of the reference: sum of gathered codes == quantized_out."""
    g = Golden("rvq_separate_fp32")
    states = g.states("s0_pre")
    q, ind, _, _ = O.rvq_forward(g["s0_x"], "fp32", states, g.cfg, training=False)
    out = O.rvq_output_from_indices([s.embed for s in states], ind)
    np.testing.assert_allclose(out, q, atol=1e-5)


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("vq_") and "heads" not in n])
def test_torch_port_is_bit_identical_to_reference(name):
    """oracle/vq_oracle_torch.py (the CPU-baseline arm of bench.py) replays the reference's ATen ops."""
    import torch
    from oracle import vq_oracle_torch as T
    g = Golden(name)
    m = g.meta
    st = T.State(torch.from_numpy(g["s0_pre_cb0_embed"]))
    st.embed_avg = torch.from_numpy(g["s0_pre_cb0_embed_avg"])[None].clone()
    st.cluster_size = torch.from_numpy(g["s0_pre_cb0_cluster_size"])[None].clone()
    for step, mode in enumerate(m["steps"]):
        x = torch.from_numpy(g[f"s{step}_x"]).to(torch.bfloat16 if m["dtype"] == "bf16" else torch.float32)
        q, i, l = T.vq_forward(x, st, cosine=m.get("use_cosine_sim", False), training=mode == "train",
                               decay=m.get("decay", 0.8), eps=m.get("eps", 1e-5),
                               commitment_weight=m.get("commitment_weight", 1.0))
        assert np.array_equal(i.numpy(), g[f"s{step}_indices"])
        np.testing.assert_allclose(q.float().numpy(), g[f"s{step}_quantize"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(float(l), float(g[f"s{step}_loss"]), rtol=1e-6)
        np.testing.assert_allclose(st.embed[0].numpy(), g[f"s{step}_post_cb0_embed"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", simvq_golden_names())
def test_simvq_oracle_matches_reference(name):
    """sim_vq.py:99-139 restated in numpy against the reference's own outputs (oracle/gen_golden.py --simvq)."""
    import os
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    q, ind, loss = O.simvq_forward(g["s0_x"], g["frozen"], g["weight"])
    assert np.array_equal(ind, g["s0_indices"])
    np.testing.assert_allclose(q, g["s0_quantize"], rtol=1e-5, atol=1e-5)   # the estimators reproduce the quantized value
    np.testing.assert_allclose(loss, g["s0_loss"], rtol=1e-5)


@pytest.mark.parametrize("name", mask_golden_names())
def test_masked_oracle_matches_reference(name):
    """`mask` / `lens` calls (vqp:1116-1119): every row is searched, the statistics see the unmasked rows only (vqp:599-600), the
    loss is the mean over the unmasked elements against the original input (vqp:1317-1325), padding comes back as zeros / the
    input with index -1 (vqp:1378-1396).  ResidualVQ hands the mask to every layer (rvq:495), the grouped module to every group."""
    g = Golden(name)
    m = g.meta
    states = g.states("s0_pre")
    vtol = 1e-5 if m["dtype"] == "fp32" else 8e-3
    for step, mode in enumerate(m["steps"]):
        mask = g[f"s{step}_mask"]
        kw = dict(training=mode == "train", mask=mask)
        if m["kind"] == "vq":
            q, ind, loss, _ = O.vq_forward(g[f"s{step}_x"], m["dtype"], states, g.cfg,
                                           return_zeros_for_masked_padding=m.get("return_zeros_for_masked_padding", True), **kw)
            assert (ind[~mask] == -1).all()
        elif m["kind"] == "rvq":
            q, ind, loss, _ = O.rvq_forward(g[f"s{step}_x"], m["dtype"], states, g.cfg, shared_codebook=m["shared_codebook"], **kw)
            assert (ind[~mask] == -1).all()
        else:
            q, ind, loss, _ = O.grouped_rvq_forward(g[f"s{step}_x"], m["dtype"], states, g.cfg, shared_codebook=m["shared_codebook"], **kw)
            assert (ind[:, ~mask] == -1).all()
        assert np.array_equal(ind, g[f"s{step}_indices"]), f"{name} step {step}"
        np.testing.assert_allclose(q, g[f"s{step}_quantize"], rtol=vtol, atol=vtol)
        np.testing.assert_allclose(loss, g[f"s{step}_loss"], rtol=1e-5 if m["dtype"] == "fp32" else 8e-3, atol=1e-7)
        for i, st in enumerate(g.flat_states(states)):
            ref = g.state(f"s{step}_post", i)
            np.testing.assert_allclose(st.cluster_size, ref.cluster_size, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(st.embed_avg, ref.embed_avg, rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(st.embed, ref.embed, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", layout_golden_names())
def test_layout_oracle_matches_reference(name):
    """Feature-map / channel-first / single-token layouts (vqp:1121-1147, restored at :1265-1277, :1364-1376)."""
    g = Golden(name)
    m = g.meta
    state = g.states("s0_pre")
    vtol = 1e-5 if m["dtype"] == "fp32" else 8e-3
    for step, mode in enumerate(m["steps"]):
        q, ind, loss, _ = O.vq_forward_layout(g[f"s{step}_x"], m["dtype"], state, g.cfg, layout=m["layout"], training=mode == "train")
        assert ind.shape == g[f"s{step}_indices"].shape and q.shape == g[f"s{step}_quantize"].shape
        assert np.array_equal(ind, g[f"s{step}_indices"]), f"{name} step {step}"
        np.testing.assert_allclose(q, g[f"s{step}_quantize"], rtol=vtol, atol=vtol)
        np.testing.assert_allclose(loss, g[f"s{step}_loss"], rtol=1e-5 if m["dtype"] == "fp32" else 8e-3, atol=1e-7)
        ref = g.state(f"s{step}_post", 0)
        np.testing.assert_allclose(state.embed, ref.embed, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", dropout_golden_names())
def test_quantize_dropout_oracle_matches_reference(name):
    """quantize_dropout (rvq:423-439, :473-476): python's random.Random(seed) picks the last active layer of a training step; the
    layers after it return index -1 / loss 0 and leave their codebooks alone."""
    g = Golden(name)
    m = g.meta
    states = g.states("s0_pre")
    Q = m["num_quantizers"]
    vtol = 1e-5 if m["dtype"] == "fp32" else 8e-3
    for step, mode in enumerate(m["steps"]):
        di = O.quantize_dropout_index(m["seeds"][step], Q, m.get("quantize_dropout_cutoff_index", 0), m.get("quantize_dropout_multiple_of", 1))
        q, ind, loss, _ = O.rvq_forward(g[f"s{step}_x"], m["dtype"], states, g.cfg, shared_codebook=m["shared_codebook"],
                                        training=mode == "train", dropout_index=di)
        assert np.array_equal(ind, g[f"s{step}_indices"]), f"{name} step {step}"
        if mode == "train":
            assert (ind[..., min(di, Q - 1) + 1:] == -1).all() and (ind[..., :di + 1] >= 0).all()
        np.testing.assert_allclose(q, g[f"s{step}_quantize"], rtol=vtol, atol=vtol)
        np.testing.assert_allclose(loss, g[f"s{step}_loss"], rtol=1e-5 if m["dtype"] == "fp32" else 8e-3, atol=1e-7)
        for i, st in enumerate(g.flat_states(states)):
            ref = g.state(f"s{step}_post", i)
            np.testing.assert_allclose(st.cluster_size, ref.cluster_size, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(st.embed, ref.embed, rtol=1e-5, atol=1e-5)
