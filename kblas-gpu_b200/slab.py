"""Contiguous batch-slab partition of a uniform batch across the GPUs of one box.

The batch is embarrassingly parallel (every matrix and its right-hand sides are independent), so
multi-GPU = one process per GPU, each owning a contiguous slab, no data-path collective.
Mirrors the reference test harness (testing/batch_triangular/test_Xpotrf_batch.cpp:108,123-131:
batchCount_gpu = batchCount / ngpu, host offset per device) but keeps the remainder instead of
silently dropping it when ngpu does not divide batchCount.
"""
from __future__ import annotations


def slab_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of the matrices owned by `rank`: ceil-sized slabs, the last ones may be shorter/empty"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-batch // world)
    b = min(batch, rank * per)
    e = min(batch, b + per)
    return b, e


def slab_offsets(batch: int, world: int, rank: int, strideA: int, elem_size: int) -> tuple[int, int, int]:
    """(first matrix, matrix count, byte offset of the slab in a strided batch)"""
    b, e = slab_range(batch, world, rank)
    return b, e - b, b * strideA * elem_size
