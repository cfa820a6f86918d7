"""Order- and shard-independent checksum of a volumetric map (host side, NumPy).

Used by bench.py (`map_chunks`, `map_hash` in every JSON line, so that runs at 1/2/4/8 GPUs and the
CPU reference arm can be compared from the driver's records alone) and by the parity tests at
sizes where holding two full voxel dumps is wasteful.

  chunk hash = position-weighted sum of the chunk's 8 KiB record (sdf | weight | colour bits, as
               1024 little-endian u64 words times fixed odd multipliers, mod 2^64), mixed with its id
  map hash   = sum of the chunk hashes mod 2^64  (a set property: independent of listing order and
               of how the chunks are split over ranks)
Works on anything that offers list_chunks() and download_chunks(ids) -> (sdf, weight, colour):
capi.Map, oracle.OracleMap.
"""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def _weights() -> np.ndarray:
    w = np.empty(1024, np.uint64)
    x = 0x9E3779B97F4A7C15
    for i in range(1024):  # splitmix64 stream, forced odd
        x = (x + 0x9E3779B97F4A7C15) & _MASK
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        w[i] = (z ^ (z >> 31)) | 1
    return w


_W = _weights()


def chunk_hashes(ids, sdf, weight, color) -> np.ndarray:
    """uint64 hash per chunk from its id (n x 3 int32) and voxel planes (n x 512 f32, n x 512 f32, n x 2048 u16)."""
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    n = len(ids)
    rec = np.concatenate([np.ascontiguousarray(sdf, np.float32).reshape(n, 512).view(np.uint64),
                          np.ascontiguousarray(weight, np.float32).reshape(n, 512).view(np.uint64),
                          np.ascontiguousarray(color, np.uint16).reshape(n, 2048).view(np.uint64)], axis=1)
    with np.errstate(over="ignore"):
        h = (rec * _W[None, :]).sum(axis=1, dtype=np.uint64)
        idh = (ids[:, 0].astype(np.int64).astype(np.uint64) * np.uint64(73856093)) ^ \
              (ids[:, 1].astype(np.int64).astype(np.uint64) * np.uint64(19349663)) ^ \
              (ids[:, 2].astype(np.int64).astype(np.uint64) * np.uint64(83492791))
        h = (h ^ idh) * np.uint64(0xD6E8FEB86659FD93)
        h ^= h >> np.uint64(32)
    return h


def map_hash(m, ids=None, slab: int = 16384):
    """(chunk count, map hash) of everything in `m` (or of the given ids), downloaded slab by slab."""
    if ids is None:
        ids = m.list_chunks()
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
    total = 0
    for a in range(0, len(ids), slab):
        part = ids[a:a + slab]
        sdf, w, col = m.download_chunks(part)
        total = (total + int(chunk_hashes(part, sdf, w, col).sum(dtype=np.uint64))) & _MASK
    return len(ids), total


def sorted_chunk_hashes(m, slab: int = 16384):
    """(ids sorted lexicographically, their chunk hashes): pin-points WHICH chunks differ."""
    ids = np.ascontiguousarray(m.list_chunks(), np.int32).reshape(-1, 3)
    ids = ids[np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))]
    hs = np.empty(len(ids), np.uint64)
    for a in range(0, len(ids), slab):
        sdf, w, col = m.download_chunks(ids[a:a + slab])
        hs[a:a + slab] = chunk_hashes(ids[a:a + slab], sdf, w, col)
    return ids, hs
