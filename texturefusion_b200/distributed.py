"""torch.distributed plumbing for the sharded path: one frame broadcast per frame.

Backend-agnostic (NCCL over NVLink on the GPU box, gloo in the CPU tests): the ingest rank
holds the frame's planes, every other rank receives them into its own frame store.  Nothing
else is exchanged per frame; per-rank results are merged on demand (sharding.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class _DevPlane:
    """CUDA array interface over a raw plane of a tf_map's frame store."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def device_plane(ptr: int, n_words: int, device: int) -> torch.Tensor:
    return torch.as_tensor(_DevPlane(ptr, n_words), device=f"cuda:{device}")


def frame_header(frame_index: int, is_keyframe: bool, pose) -> torch.Tensor:
    """[frame_index, is_keyframe, pose(16)] as one float64 tensor (exact for int32 indices)."""
    h = np.zeros(18, np.float64)
    h[0], h[1] = frame_index, 1.0 if is_keyframe else 0.0
    h[2:] = np.asarray(pose, np.float64).reshape(16)
    return torch.from_numpy(h)


def broadcast_frame(planes, header: torch.Tensor | None = None, src: int = 0, group=None):
    """Broadcast the frame's planes (depth [, rgba, quality]) — device tensors under NCCL,
    CPU tensors under gloo — and optionally the small header.  Returns the header."""
    if header is not None:
        if planes and planes[0].is_cuda:
            header = header.to(planes[0].device)
        dist.broadcast(header, src=src, group=group)
    for p in planes:
        dist.broadcast(p, src=src, group=group)
    return header


def gather_lists(ids: np.ndarray, payloads=(), dst: int = 0, group=None):
    """On-demand gather of per-rank chunk lists (+ parallel arrays) to `dst` (host side)."""
    world = dist.get_world_size(group)
    obj = (np.asarray(ids, np.int32), tuple(np.asarray(p) for p in payloads))
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(obj, out, dst=dst, group=group)
    return out
