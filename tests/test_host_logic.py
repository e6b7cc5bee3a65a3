"""Host-side logic that needs no GPU: chunk schedule of the host-buffer pipeline, packed statistics layout."""
import pytest

from vector_quantize_pytorch_b200 import dist as vdist
from vector_quantize_pytorch_b200.vector_quantize import host_chunk_bounds


@pytest.mark.parametrize("N", [1, 100, 255, 256, 257, 1000, 4096, 65536, 262144, 262145, 1000003])
@pytest.mark.parametrize("n_chunks", [1, 2, 3, 8, 10, 16, 64])
def test_host_chunk_bounds_cover_the_batch(N, n_chunks):
    b = host_chunk_bounds(N, n_chunks)
    assert b[0] == 0 and b[-1] == N
    assert all(lo < hi for lo, hi in zip(b, b[1:])), "empty or reversed chunk"
    assert len(b) - 1 <= n_chunks
    assert all(x % 256 == 0 for x in b[1:-1]), "interior boundaries must be whole CTA-pair tiles"
    if N >= 256 * 3 * n_chunks:  # enough rows for the ramp to show: outer chunks are the short ones
        sizes = [hi - lo for lo, hi in zip(b, b[1:])]
        assert sizes[0] <= max(sizes) and sizes[-1] <= max(sizes)
        if n_chunks >= 5:
            assert sizes[0] * 2 <= max(sizes) + 512 and sizes[-1] * 2 <= max(sizes) + 512


def test_stats_layout_is_16_byte_aligned():
    # [cluster_size padded to a multiple of 4 | embed_sum (K, D)] per codebook, back to back: every slice and every
    # embed_sum block starts on a 16-byte boundary (vector REDs / float4 accesses in the kernels)
    for K, D in ((5, 8), (48, 32), (1024, 256), (16384, 512)):
        offsets, sizes, total = vdist.stats_layout([(K, D)] * 3)
        assert offsets[0] == 0 and total == sum(sizes)
        assert all(o % 4 == 0 for o in offsets) and all(s % 4 == 0 for s in sizes)
        assert sizes[0] >= K + K * D and sizes[0] - (K + K * D) < 4


def test_shard_rows_partitions_the_batch():
    for n in (1, 7, 262144, 262145):
        for world in (1, 2, 3, 8):
            spans = [vdist.shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            lens = [hi - lo for lo, hi in spans]
            assert max(lens) - min(lens) <= 1


def test_plan_cache_is_never_copied_with_the_module():
    """The cached op lists of ResidualVQ / GroupedResidualVQ hold raw device pointers: deepcopy / pickle of a module must
    start with an empty cache instead of copying them."""
    import copy, pickle
    from vector_quantize_pytorch_b200.residual_vq import _PlanCache
    c = _PlanCache()
    c[("key",)] = (object(), object())
    assert len(copy.deepcopy(c)) == 0
    assert len(pickle.loads(pickle.dumps(c))) == 0
    holder = {"_plans": c, "other": [1, 2]}
    dup = copy.deepcopy(holder)
    assert dup["other"] == [1, 2] and len(dup["_plans"]) == 0 and isinstance(dup["_plans"], _PlanCache)


def test_rvq_op_layout_matches_the_header():
    """ctypes mirror of vqb_rvq_op (include/vqb200.h): the nested structs must sit where the C compiler puts them."""
    import ctypes, subprocess, tempfile, os, shutil
    from vector_quantize_pytorch_b200 import _C
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        import pytest
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "vqb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu", sizeof(vqb_rvq_op), ' \
          'offsetof(vqb_rvq_op, stage), offsetof(vqb_rvq_op, ema), offsetof(vqb_rvq_op, acc), sizeof(vqb_vq_forward_args), ' \
          'offsetof(vqb_rvq_op, bar), offsetof(vqb_rvq_op, emap), offsetof(vqb_rvq_op, emap.scratch));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run([cc, "-I", os.path.join(root, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        got = [int(v) for v in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(_C.RvqOp), _C.RvqOp.stage.offset, _C.RvqOp.ema.offset, _C.RvqOp.acc.offset,
                   ctypes.sizeof(_C.VQForwardArgs), _C.RvqOp.bar.offset, _C.RvqOp.emap.offset,
                   _C.RvqOp.emap.offset + _C.RvqEmaPeersArgs.scratch.offset]


def test_quantize_dropout_layer_count_follows_the_reference():
    """ResidualVQ._active_layers (host logic, no GPU needed): python's random.Random(seed).randrange(cutoff, Q) is the last
    active layer, rounded up to a multiple (rvq:434-439); eval mode and quantize_dropout=False run every layer; Q == 1 switches
    dropout off (rvq:253)."""
    import math
    import random
    import vector_quantize_pytorch_b200 as m
    from oracle import vq_oracle as O
    for Q, cutoff, mult in [(4, 0, 1), (6, 1, 2), (8, 2, 4), (8, 0, 3), (5, 4, 1)]:
        rvq = m.ResidualVQ(dim=16, num_quantizers=Q, codebook_size=8, quantize_dropout=True, quantize_dropout_cutoff_index=cutoff,
                           quantize_dropout_multiple_of=mult)
        rvq.train()
        seen = set()
        for seed in range(200):
            idx = random.Random(seed).randrange(cutoff, Q)            # rvq:436
            if mult != 1:
                idx = math.ceil((idx + 1) / mult) * mult - 1          # rvq:439
            assert idx == O.quantize_dropout_index(seed, Q, cutoff, mult)
            n = rvq._active_layers(seed, "cpu")
            assert n == min(idx + 1, Q) and cutoff + 1 <= n <= Q
            seen.add(n)
        assert len(seen) > 1 or cutoff == Q - 1
        rvq.eval()
        assert rvq._active_layers(3, "cpu") == Q                        # no dropout outside training (rvq:423)
    assert m.ResidualVQ(dim=16, num_quantizers=3, codebook_size=8).train()._active_layers(3, "cpu") == 3
    assert not m.ResidualVQ(dim=16, num_quantizers=1, codebook_size=8, quantize_dropout=True).quantize_dropout


def test_coarse_indices_are_padded_with_dropped_layers():
    """rvq:333-339: fewer than num_quantizers index columns decode as if the missing layers had been dropped (-1)."""
    import torch
    import vector_quantize_pytorch_b200 as m
    rvq = m.ResidualVQ(dim=16, num_quantizers=4, codebook_size=8, quantize_dropout=True)
    ind = torch.randint(0, 8, (2, 5, 2))
    padded = rvq._pad_dropped(ind)
    assert padded.shape == (2, 5, 4) and torch.equal(padded[..., :2], ind) and (padded[..., 2:] == -1).all()
    assert rvq._pad_dropped(padded) is padded
    import pytest
    with pytest.raises(AssertionError):
        m.ResidualVQ(dim=16, num_quantizers=4, codebook_size=8)._pad_dropped(ind)   # rvq:336: only with quantize dropout
