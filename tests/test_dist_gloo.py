"""N>1 host logic on CPU: world_size-2 gloo run of the packed-statistics exchange.

Each rank owns a shard of the batch, computes its shard's statistics with the oracle (the CUDA kernels
need a GPU; their output is checked against the same oracle in tests/test_parity_gpu.py), packs them
with the PRODUCT's layout + all-reduce helpers, and applies the EMA.  Both ranks must end with the same
codebook as the single-process oracle run over the whole batch — the reference semantics of
vector_quantize_pytorch.py:603-617.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vector_quantize_pytorch_b200 import dist as vd
    rng = np.random.default_rng(7)
    books = [(40, 16), (5, 8), (64, 32)]  # (K, D) incl. a K that is not a multiple of 4
    N = 301
    offsets, sizes, total = vd.stats_layout(books)
    packed = torch.zeros(total, dtype=torch.float32)
    xs, idxs, embeds = [], [], []
    for (K, D), off, size in zip(books, offsets, sizes):
        x = rng.standard_normal((N, D), dtype=np.float32)
        e = rng.standard_normal((K, D), dtype=np.float32)
        idx = O.argmax_first(O.scores(x, e, False))
        xs.append(x); idxs.append(idx); embeds.append(e)
        a, b = vd.shard_rows(N, world, rank)
        cs, es = O.batch_stats(x[a:b], idx[a:b], K)
        sl_cs, sl_es = vd.split_stats(packed[off:off + size], K, D)
        sl_cs.copy_(torch.from_numpy(cs))
        sl_es.copy_(torch.from_numpy(es))
    assert vd.is_distributed()
    vd.allreduce_packed(packed)
    result = []
    for (K, D), off, size, x, idx, e in zip(books, offsets, sizes, xs, idxs, embeds):
        cs, es = vd.split_stats(packed[off:off + size], K, D)
        cs_ref, es_ref = O.batch_stats(x, idx, K)  # whole batch in one process
        np.testing.assert_array_equal(cs.numpy(), cs_ref)
        np.testing.assert_allclose(es.numpy(), es_ref, rtol=1e-5, atol=1e-5)
        st = O.CodebookState.from_embed(e)
        O.ema_inplace(st.cluster_size, cs.numpy(), 0.8)
        O.ema_inplace(st.embed_avg, es.numpy(), 0.8)
        O.update_ema(st, 1e-5, False)
        ref = O.CodebookState.from_embed(e)
        O.track_stats(ref, x, idx, 0.8)
        O.update_ema(ref, 1e-5, False)
        np.testing.assert_allclose(st.embed, ref.embed, rtol=1e-5, atol=1e-6)
        result.append(st.embed)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([r.ravel() for r in result]))
    dist.destroy_process_group()


def test_packed_stats_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b), "replicas diverged"


def test_shard_rows_cover_the_batch():
    from vector_quantize_pytorch_b200 import dist as vd
    for n in (1, 7, 128, 1000):
        for w in (1, 2, 3, 8):
            edges = [vd.shard_rows(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
