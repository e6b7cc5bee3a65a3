"""Helpers to replay tests/golden/*.npz (written by oracle/gen_golden.py from the real reference)."""
import glob
import json
import os

import numpy as np

from oracle import vq_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Small fixtures that store their inputs (the `big*` ones are replayed by tests/test_big_golden.py)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith(("big", "grad", "simvq", "mask", "layout", "dropout")))


def simvq_golden_names():
    """SimVQ fixtures (oracle/gen_golden.py --simvq): inputs, codebook, transform weight, the reference's outputs and gradients."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "simvq_*.npz"))))


def mask_golden_names():
    """Variable-length fixtures (oracle/gen_golden.py --mask): `mask` / `lens` calls of the reference, s{step}_mask stored."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "mask_*.npz"))))


def layout_golden_names():
    """Input-layout fixtures (oracle/gen_golden.py --layout): feature maps, channel-first, one token per batch element."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "layout_*.npz"))))


def dropout_golden_names():
    """quantize_dropout fixtures (oracle/gen_golden.py --dropout): one explicit dropout seed per step in meta["seeds"]."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "dropout_*.npz"))))


def grad_golden_names():
    """Gradient fixtures (oracle/gen_golden.py --grad): x, upstream gradient G, the reference's d/dx."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "grad_*.npz"))))


def replayable_on_gpu(names):
    """Fixtures whose outputs do not depend on the CPU RNG (dead-code expiry draws torch.randperm on the tensor's device)."""
    return [n for n in names if not n.startswith(("expire", "kmeans", "grad"))]


def cpu_pick_fn(n, num):
    """sample_vectors (vqp:156-163) on torch's global CPU generator — the caller seeds it like oracle/gen_golden.py."""
    import torch
    return (torch.randperm(n)[:num] if n >= num else torch.randint(0, n, (num,))).numpy()


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.name = name

    def __getitem__(self, k):
        return self.z[k]

    @property
    def cfg(self):
        m = self.meta
        dim = m.get("codebook_dim", m["dim"] // m.get("groups", 1))
        return O.VQConfig(dim=dim, codebook_size=m["codebook_size"], use_cosine_sim=m.get("use_cosine_sim", False), heads=m.get("heads", 1),
                          separate_codebook_per_head=m.get("separate_codebook_per_head", False),
                          decay=m.get("decay", 0.8), eps=m.get("eps", 1e-5),
                          commitment_weight=m.get("commitment_weight", 1.0),
                          threshold_ema_dead_code=m.get("threshold_ema_dead_code", 0), kmeans_iters=m.get("kmeans_iters", 10))

    def state(self, tag, i):
        return O.CodebookState(self.z[f"{tag}_cb{i}_embed"].copy(), self.z[f"{tag}_cb{i}_embed_avg"].copy(),
                               self.z[f"{tag}_cb{i}_cluster_size"].copy(),
                               initted=not (self.meta.get("kmeans_init", False) and tag == "s0_pre"))

    def states(self, tag):
        """Codebook states arranged as the oracle entry points expect them."""
        m = self.meta
        n = m["n_codebooks"]
        flat = [self.state(tag, i) for i in range(n)]
        if m["kind"] == "vq":
            return flat if m.get("separate_codebook_per_head") else flat[0]
        Q = m["num_quantizers"]
        if m["kind"] == "rvq":
            return [flat[0]] * Q if m["shared_codebook"] else flat
        G = m["groups"]
        if m["shared_codebook"]:
            return [[flat[g]] * Q for g in range(G)]
        return [flat[g * Q:(g + 1) * Q] for g in range(G)]

    def flat_states(self, states):
        m = self.meta
        if m["kind"] == "vq":
            return list(states) if m.get("separate_codebook_per_head") else [states]
        if m["kind"] == "rvq":
            return [states[0]] if m["shared_codebook"] else list(states)
        out = []
        for g in states:
            out += [g[0]] if m["shared_codebook"] else list(g)
        return out


def near_tie_rows(x, embed, cosine, tol=2e-6):
    """Rows whose two best float64 scores are within fp32 rounding noise of each other."""
    _, gap = O.top2_gap_f64(x, embed, cosine)
    return gap < tol
