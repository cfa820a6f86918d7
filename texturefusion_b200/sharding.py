"""Chunk ownership and per-rank result merging for the chunk-sharded multi-GPU path.

One process per GPU; rank r owns the chunks whose 8x8x8-chunk block hashes to r with the
reference's ChunkHasher (Structure/ChunkManager.h:44-53), folded to 32 bits.  The device applies
the same function inside cull_kernel (tf_device.cuh: owner_of); this module is the host-side
mirror used to route host-provided chunk lists and to merge per-rank results back into the
reference's traversal order.  There is no per-frame collective besides the frame broadcast.
"""
from __future__ import annotations

import numpy as np

_P1, _P2, _P3 = 73856093, 19349663, 83492791
OWNER_SHIFT = 3  # == kOwnerShift in csrc/tf_device.cuh


def owner_of(ids, n_ranks: int) -> np.ndarray:
    """Rank owning each chunk id (N x 3 int32)."""
    ids = np.asarray(ids, np.int64).reshape(-1, 3)
    b = ids >> OWNER_SHIFT  # floor(id / 8): the owner block the chunk lies in
    # two's complement wrap-around of size_t arithmetic, as the device computes it
    with np.errstate(over="ignore"):
        h = (b[:, 0].astype(np.uint64) * np.uint64(_P1)) ^ (b[:, 1].astype(np.uint64) * np.uint64(_P2)) ^ \
            (b[:, 2].astype(np.uint64) * np.uint64(_P3))
    h32 = (h & np.uint64(0xffffffff)) ^ (h >> np.uint64(32))
    return (h32 % np.uint64(n_ranks)).astype(np.int32)


def split_by_owner(ids, n_ranks: int):
    """Partition a chunk list by owner, keeping the order inside each part."""
    ids = np.asarray(ids, np.int32).reshape(-1, 3)
    own = owner_of(ids, n_ranks)
    return [ids[own == r] for r in range(n_ranks)], own


def traversal_order(ids, min_id, step: int) -> np.ndarray:
    """Permutation that puts `ids` into GetChunkIDsObservedByCamera's emission order
    (Structure/ChunkManager.h:472-545): coarse block (x, y, z) from min_id-1 in strides of
    `step`, then the children (i, j, k) inside the block."""
    ids = np.asarray(ids, np.int64).reshape(-1, 3)
    rel = ids - (np.asarray(min_id, np.int64) - 1)
    blk, sub = rel // step, rel % step
    return np.lexsort((sub[:, 2], sub[:, 1], sub[:, 0], blk[:, 2], blk[:, 1], blk[:, 0]))


def merge_rank_lists(rank_ids, rank_payloads=None, *, min_id=None, step=None):
    """Concatenate per-rank chunk lists (and parallel payload arrays) into one list.
    With min_id/step the result is in the reference's traversal order; otherwise it is
    sorted lexicographically (the caller consumes it as a set, e.g. kf.validChunks)."""
    ids = np.concatenate([np.asarray(a, np.int32).reshape(-1, 3) for a in rank_ids], axis=0)
    if min_id is not None and step is not None:
        order = traversal_order(ids, min_id, step)
    else:
        order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
    out = [ids[order]]
    if rank_payloads is not None:
        for cols in zip(*rank_payloads):
            out.append(np.concatenate(cols, axis=0)[order])
    return out[0] if rank_payloads is None else tuple(out)
