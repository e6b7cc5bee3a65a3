"""Generate the BASELINE-shape fixtures tests/golden/big_*.npz by running the UNMODIFIED reference
(TEST INFRASTRUCTURE ONLY; needs /root/reference, i.e. runs in the build container):

    python oracle/gen_golden_big.py [case ...]

Inputs and initial codebooks are NOT stored: they are regenerated from the seeded recipe in oracle/big_cases.py (the
fixture keeps checksums of them).  Stored per step: the reference's indices (int16), loss(es), the first rows and the
column sums of `quantize`, and of every codebook's post-step state the full cluster_size, sampled rows and column
sums of embed / embed_avg.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402
import big_cases as B  # noqa: E402
from gen_golden import codebooks_of, f32  # noqa: E402


def make_warm_state(ref):
    """Default-init codebook (vqp:112-115 distribution) after ONE reference training step on a seeded fp32 batch:
    the 'warm' regime every later step of a real run is in."""
    gen = torch.Generator().manual_seed(4242)
    vq = ref.VectorQuantize(dim=256, codebook_size=1024)
    cb = codebooks_of(vq)[0]
    e = B.default_init(1024, 256, gen)
    cb.embed.data.copy_(e[None]); cb.embed_avg.data.copy_(e[None])
    x = torch.randn(8, 4096, 256, generator=gen)
    vq.train()
    with torch.no_grad():
        vq(x)
    np.savez_compressed(B.WARM_STATE, embed=f32(cb.embed[0]), embed_avg=f32(cb.embed_avg[0]),
                        cluster_size=f32(cb.cluster_size[0]))
    print(f"warm state: {os.path.getsize(B.WARM_STATE) / 1024:.0f} KiB, codes used {(cb.cluster_size[0] > 0.9).sum().item()}")


def build(ref, case):
    kw = case["kw"]
    if case["kind"] == "vq":
        return ref.VectorQuantize(**kw)
    if case["kind"] == "rvq":
        return ref.ResidualVQ(**kw)
    return ref.GroupedResidualVQ(**kw)


def run_case(ref, name):
    case = B.CASES[name]
    torch.manual_seed(1234)
    module = build(ref, case)
    books = codebooks_of(module)
    inits = B.initial_states(case)
    assert len(books) == len(inits), (len(books), len(inits))
    store = {}
    for cb, (e, ea, cs) in zip(books, inits):
        cb.embed.data.copy_(e[None]); cb.embed_avg.data.copy_(ea[None]); cb.cluster_size.data.copy_(cs[None])
    meta = dict(name=name, kind=case["kind"], kw=case["kw"], shape=list(case["shape"]), dtype=case["dtype"], init=case["init"],
                seed=case["seed"], steps=case["steps"], torch=torch.__version__, n_codebooks=len(books),
                init_digest=[B.digest(e) for e, _, _ in inits], x_digest=[])
    module.train()
    for s, x in enumerate(B.step_inputs(case)):
        meta["x_digest"].append(B.digest(x))
        with torch.no_grad():
            q, ind, loss = module(x)[:3]
        K = case["kw"]["codebook_size"]
        store[f"s{s}_indices"] = ind.cpu().numpy().astype(np.int16 if K <= 32767 else np.int32)
        store[f"s{s}_loss"] = f32(loss)
        qf = q.float().reshape(-1, q.shape[-1])
        store[f"s{s}_q_rows"] = f32(qf[:64])
        store[f"s{s}_q_colsum"] = qf.double().sum(0).numpy()
        for i, cb in enumerate(books):
            rows = B.sample_rows(K)
            store[f"s{s}_cb{i}_cluster_size"] = f32(cb.cluster_size[0])
            for nm, t in (("embed", cb.embed[0]), ("embed_avg", cb.embed_avg[0])):
                store[f"s{s}_cb{i}_{nm}_rows"] = f32(t[rows])
                store[f"s{s}_cb{i}_{nm}_colsum"] = t.double().sum(0).numpy()
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(B.GOLDEN, name + ".npz")
    np.savez_compressed(path, **store)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or list(B.CASES)
    if not os.path.exists(B.WARM_STATE) or "--warm" in names:
        make_warm_state(ref)
    for n in names:
        if n in B.CASES:
            run_case(ref, n)


if __name__ == "__main__":
    main()
