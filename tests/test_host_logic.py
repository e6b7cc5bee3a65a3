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
