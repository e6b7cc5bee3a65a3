"""Multi-GPU plumbing of the EMA update.

The path shards naturally: tokens are independent given the codebook, so the batch is split over the
ranks (one process per GPU) and the only exchange is the SUM of the per-rank batch statistics.  The
reference issues two all-reduces per codebook per stage (vector_quantize_pytorch.py:603, :607); here
every codebook touched by a forward writes its statistics into ONE packed fp32 buffer
`[cluster_size (K, padded to 4) | embed_sum (K x D)]*` and a single `all_reduce` (NCCL over NVLink on
B200, gloo in the CPU tests) covers them all.  After the reduction every rank applies the identical EMA
kernel to identical numbers, so the replicas' codebooks stay bit-identical.
"""
from __future__ import annotations

import torch
import torch.distributed as distributed

from ._C import lib


def is_distributed() -> bool:
    return distributed.is_available() and distributed.is_initialized() and distributed.get_world_size() > 1


def stats_layout(codebooks: list[tuple[int, int]]) -> tuple[list[int], list[int], int]:
    """codebooks: [(K, D), ...] -> (offsets, sizes, total) in floats of the packed statistics buffer."""
    sizes = [int(lib.vqb_stats_floats(K, D)) for K, D in codebooks]
    offsets, pos = [], 0
    for s in sizes:
        offsets.append(pos)
        pos += s
    return offsets, sizes, pos


def split_stats(stats: torch.Tensor, K: int, D: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Views (cluster_size (K,), embed_sum (K, D)) into one codebook's slice of the packed buffer."""
    off = int(lib.vqb_stats_offset(K))
    return stats[:K], stats[off:off + K * D].view(K, D)


def allreduce_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM over ranks of the packed statistics (no-op outside a process group)."""
    if is_distributed():
        distributed.all_reduce(packed, op=distributed.ReduceOp.SUM, group=group)
    return packed


def shard_rows(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of the batch rows over the ranks."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
