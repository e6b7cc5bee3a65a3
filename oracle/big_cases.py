"""Seeded recipes of the BASELINE-shape parity fixtures  —  TEST INFRASTRUCTURE ONLY.

`oracle/gen_golden_big.py` runs the UNMODIFIED reference on these inputs and commits its outputs under
`tests/golden/big_*.npz`; the tests regenerate the very same inputs from the recipe (torch's CPU generator is
deterministic for a given torch build; every fixture carries checksums of its inputs so that a drift is detected,
not silently compared).  Storing the inputs themselves would cost 16-130 MB per fixture.

Shapes follow BASELINE.json configs[1..4] with fewer rows:
  cfg2  VectorQuantize   D=256 K=1024, 32768 rows, bf16 / fp32, warm (post-EMA) and default-init codebook
  cfg3  ResidualVQ       D=256 K=1024 Q=8 shared codebook, 8192 rows, bf16 / fp32, warm codebook
  cfg4  VectorQuantize   D=512 K=16384 cosine, 4096 rows, bf16 / fp32
  cfg5  GroupedResidualVQ dim=256 G=2 Q=8 K=1024, 8192 rows, fp32
"""
from __future__ import annotations

import hashlib
import math
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
WARM_STATE = os.path.join(GOLDEN, "bigstate_warm_k1024_d256.npz")

CASES = {
    # name: kind, module kwargs, x shape, dtype, init, seed, steps
    "big_cfg2_bf16_warm": dict(kind="vq", kw=dict(dim=256, codebook_size=1024), shape=(8, 4096, 256), dtype="bf16",
                               init="warm", seed=101, steps=2),
    "big_cfg2_fp32_warm": dict(kind="vq", kw=dict(dim=256, codebook_size=1024), shape=(8, 4096, 256), dtype="fp32",
                               init="warm", seed=102, steps=2),
    "big_cfg2_bf16_default": dict(kind="vq", kw=dict(dim=256, codebook_size=1024), shape=(8, 4096, 256), dtype="bf16",
                                  init="default", seed=103, steps=1),
    "big_cfg2_fp32_default": dict(kind="vq", kw=dict(dim=256, codebook_size=1024), shape=(8, 4096, 256), dtype="fp32",
                                  init="default", seed=104, steps=1),
    "big_cfg3_bf16": dict(kind="rvq", kw=dict(dim=256, num_quantizers=8, codebook_size=1024, shared_codebook=True),
                          shape=(2, 4096, 256), dtype="bf16", init="warm", seed=105, steps=2),
    "big_cfg3_fp32": dict(kind="rvq", kw=dict(dim=256, num_quantizers=8, codebook_size=1024, shared_codebook=True),
                          shape=(2, 4096, 256), dtype="fp32", init="warm", seed=106, steps=2),
    "big_cfg4_bf16": dict(kind="vq", kw=dict(dim=512, codebook_size=16384, use_cosine_sim=True), shape=(1, 4096, 512),
                          dtype="bf16", init="randn", seed=107, steps=2),
    "big_cfg4_fp32": dict(kind="vq", kw=dict(dim=512, codebook_size=16384, use_cosine_sim=True), shape=(1, 4096, 512),
                          dtype="fp32", init="randn", seed=108, steps=2),
    "big_cfg5_fp32": dict(kind="grvq", kw=dict(dim=256, groups=2, num_quantizers=8, codebook_size=1024), shape=(2, 4096, 256),
                          dtype="fp32", init="randn", seed=109, steps=2),
}

TDT = {"fp32": torch.float32, "bf16": torch.bfloat16}


def n_codebooks(case) -> int:
    kw = case["kw"]
    if case["kind"] == "vq":
        return 1
    per = 1 if kw.get("shared_codebook") else kw["num_quantizers"]
    return per * kw.get("groups", 1)


def codebook_dim(case) -> int:
    return case["kw"]["dim"] // case["kw"].get("groups", 1)


def default_init(K: int, D: int, gen) -> torch.Tensor:
    """The distribution of the reference's default codebook init (vqp:112-115: kaiming_uniform_ on an (H, K, D)
    tensor -> U(-b, b), b = sqrt(6 / (K*D))), drawn from OUR seeded generator so that it can be regenerated."""
    b = math.sqrt(6.0 / (K * D))
    return (torch.rand(K, D, generator=gen) * 2 - 1) * b


def initial_states(case):
    """List of (embed, embed_avg, cluster_size) fp32 torch tensors, one per distinct codebook, forward order."""
    K, D = case["kw"]["codebook_size"], codebook_dim(case)
    gen = torch.Generator().manual_seed(case["seed"] * 7919 + 1)
    out = []
    for _ in range(n_codebooks(case)):
        if case["init"] == "warm":
            z = np.load(WARM_STATE)
            out.append((torch.from_numpy(z["embed"]).clone(), torch.from_numpy(z["embed_avg"]).clone(),
                        torch.from_numpy(z["cluster_size"]).clone()))
            continue
        if case["init"] == "default":
            e = default_init(K, D, gen)
        else:
            e = torch.randn(K, D, generator=gen)
            if case["kw"].get("use_cosine_sim"):
                e = torch.nn.functional.normalize(e, dim=-1)
        out.append((e, e.clone(), torch.ones(K)))
    return out


def step_inputs(case):
    """x of every step, in the case's dtype."""
    gen = torch.Generator().manual_seed(case["seed"])
    return [torch.randn(*case["shape"], generator=gen).to(TDT[case["dtype"]]) for _ in range(case["steps"])]


def digest(t: torch.Tensor) -> str:
    a = t.detach().contiguous().cpu()
    a = a.view(torch.int16) if a.dtype == torch.bfloat16 else a
    return hashlib.sha256(a.numpy().tobytes()).hexdigest()[:16]


def sample_rows(n: int) -> np.ndarray:
    """Row subset stored for big (K x D) tensors: the first 24 rows and 40 strided ones."""
    return np.unique(np.concatenate([np.arange(min(24, n)), np.linspace(0, n - 1, 40).astype(np.int64)]))
