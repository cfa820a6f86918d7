"""Host-side mirror of the reference's fusion interface, over the C ABI.

Same class / method names, argument meaning and error behaviour as the C++ classes that
GCFusion/MobileFusion.cpp drives (chisel::Chisel, Structure/Chisel.h:46-493;
chisel::ChunkManager, Structure/ChunkManager.h:119-207; chisel::Atlas, Structure/Atlas.h:43-75),
so that the parity tests read like calls into the reference.  The C++ twin of this file,
for linking into FlashFusion itself, is texturefusion_b200/host/chisel_b200.h.

All voxel work happens in libtexfusion_b200.so on the GPU; this file only keeps the
host-side containers the reference keeps on the host (meshesToUpdate, per-chunk
observations, validChunks lists).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi

_SCRATCH_FRAME = 0x7F000000  # frame-store index used for images passed by pointer

_NEIGHBOURS = ((0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))


@dataclass
class PinholeCamera:
    """chisel::PinholeCamera (3rd_party/open_chisel/camera/PinholeCamera.h:33-75).  The
    int-returning getters of the reference are applied inside the library."""

    fx: float = 525.0
    fy: float = 525.0
    cx: float = 319.5
    cy: float = 239.5
    width: int = 640
    height: int = 480
    near: float = 0.01
    far: float = 5.0

    def SetIntrinsics(self, fx, fy, cx, cy):
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy

    def SetNearPlane(self, v):
        self.near = v

    def SetFarPlane(self, v):
        self.far = v

    def SetWidth(self, v):
        self.width = v

    def SetHeight(self, v):
        self.height = v


@dataclass
class ChunkView:
    """What callers read from a chisel::ChunkPtr: the id and the observations map
    (GCFusion/MobileFusion.cpp:258-260, Structure/TexMap.cpp:69-83)."""

    ID: tuple
    observations: dict = field(default_factory=dict)


class ChunkManager:
    """chisel::ChunkManager facade: chunk existence, observations, voxel read-back."""

    def __init__(self, m: capi.Map, resolution: float):
        self._m = m
        self._res = resolution
        self._obs: dict[tuple, dict] = {}

    def GetResolution(self):
        return self._res

    def GetChunkSize(self):
        return (8, 8, 8)

    def HasChunk(self, chunk_id) -> bool:
        return self._m.has_chunk(chunk_id)

    def GetChunk(self, chunk_id) -> ChunkView:
        cid = tuple(int(v) for v in chunk_id)
        if not self._m.has_chunk(cid):
            raise KeyError(cid)  # unordered_map::at -> std::out_of_range
        return ChunkView(cid, self._obs.setdefault(cid, {}))

    def RemoveChunk(self, chunk_id) -> bool:
        cid = tuple(int(v) for v in chunk_id)
        had = self._m.has_chunk(cid)
        if had:
            self._m.remove_chunks(np.array([cid], np.int32))
            self._obs.pop(cid, None)
        return had

    def GetChunkCount(self) -> int:
        return self._m.chunk_count()

    def GetChunkIDs(self) -> np.ndarray:
        return self._m.list_chunks()

    def GetVoxels(self, ids):
        """chunk->voxels.sdf / weight and chunk->colors.colorData for the CPU mesher
        (Structure/ChunkManager.cpp:614-626)."""
        return self._m.download_chunks(ids)


class Atlas:
    """chisel::Atlas (Structure/Atlas.{h,cpp}): slot placement and patch copy / resize."""

    MAX_PATCH_WIDTH = 96 * 72 * 2
    MAX_PATCH_HEIGHT = 72 * 96 * 2

    def __init__(self, m: capi.Map):
        self._m = m
        self.PATCH_WIDTH, self.PATCH_HEIGHT = m.atlas_patch_size()
        self.hot_start = 0
        self.hot_end = 0
        self._patches: dict[tuple, dict] = {}

    def HasPatch(self, chunk_id) -> bool:
        return tuple(int(v) for v in chunk_id) in self._patches

    def AddPatch(self, chunk_id) -> dict:
        cid = tuple(int(v) for v in chunk_id)
        if cid not in self._patches:
            try:
                texloc = self._m.atlas_alloc_slot(cid)
            except capi.TexFusionError as e:
                if e.code == capi.TF_ERR_ATLAS_FULL:
                    raise OverflowError("No enough space for texture storage.") from e
                raise
            self._patches[cid] = {"texloc": texloc, "frameid": -1, "box": None}
        else:  # Patch::clear (Structure/Patch.cpp:177-189)
            self._patches[cid].update(frameid=-1, box=None)
        return self._patches[cid]

    def GetTexLoc(self, chunk_id):
        k = self._patches[tuple(int(v) for v in chunk_id)]["texloc"]
        return (k % self.MAX_PATCH_WIDTH, k // self.MAX_PATCH_WIDTH)

    def SetPatchImage(self, chunk_id, frame_index: int, box):
        """Patch::SetFrameid + Patch::SetImage (Structure/Patch.cpp:172-175): the crop
        `box` = (x, y, w, h) of key-frame `frame_index`'s rgb, which must be in the store."""
        p = self._patches[tuple(int(v) for v in chunk_id)]
        p["frameid"], p["box"] = frame_index, tuple(int(v) for v in box)

    def UpdateBuffers(self, chunk_ids):
        """Chisel::UpdateAtlas (Structure/Chisel.cpp:191-196) -> Atlas::UpdateBuffer for each id."""
        descs = []
        for cid in chunk_ids:
            p = self._patches.get(tuple(int(v) for v in cid))
            if p is None or p["box"] is None or p["frameid"] < 0:
                continue
            descs.append((p["texloc"], p["frameid"], *p["box"]))
        self._m.atlas_update(descs)

    def texture_rows(self, hot_start=None, hot_end=None) -> np.ndarray:
        hs = self.hot_start if hot_start is None else hot_start
        he = self.hot_end if hot_end is None else hot_end
        return self._m.atlas_download(hs, he)


class Chisel:
    """chisel::Chisel (Structure/Chisel.h:46-493), fusion part."""

    def __init__(self, chunkSize=(8, 8, 8), voxelResolution=0.005, useColor=True, *, width=640, height=480,
                 device=0, n_ranks=1, rank=0, max_chunks=0, max_frames=0, trunc=capi.DEFAULT_TRUNC):
        if tuple(chunkSize) != (8, 8, 8):
            raise ValueError("only 8x8x8 chunks (GCFusion/MobileFusion.h:231-233)")
        self.map = capi.Map(voxelResolution, use_color=useColor, trunc=trunc, device=device, n_ranks=n_ranks,
                            rank=rank, max_chunks=max_chunks, max_frames=max_frames, width=width, height=height)
        self.chunkManager = ChunkManager(self.map, float(np.float32(voxelResolution)))
        self.meshesToUpdate: dict[tuple, bool] = {}
        self.atlas = Atlas(self.map)

    def GetChunkManager(self):
        return self.chunkManager

    def Reset(self):
        self.map.reset()
        self.meshesToUpdate.clear()
        self.chunkManager._obs.clear()

    # Structure/Chisel.h:103-140
    def PrepareIntersectChunks(self, depthImage, depthExtrinsic, depthCamera):
        """Returns (chunksIntersecting, needsUpdateFlag, newChunkFlag)."""
        self.map.upload_frame(_SCRATCH_FRAME, depthImage)
        ids, new = self.map.prepare(_SCRATCH_FRAME, depthExtrinsic, depthCamera)
        return ids, np.zeros(len(ids), np.uint8), new

    # Structure/Chisel.h:218-249 (list form) and :453-468 (convenience form)
    def IntegrateDepthScanColor(self, depthImage, colorImage, depthExtrinsic, depthCamera,
                                chunksIntersecting=None, needsUpdateFlag=None, integrate_flag=1, keyframeID=-1,
                                observationQualityPointer=None):
        self.map.upload_frame(_SCRATCH_FRAME, depthImage, colorImage,
                              observationQualityPointer if colorImage is not None else None)
        use_color = colorImage is not None
        if chunksIntersecting is None:
            st, ids, new, upd, q = self.map.integrate_frame(_SCRATCH_FRAME, use_color, depthExtrinsic, depthCamera)
            self._mark(ids, upd)
            return st
        if len(chunksIntersecting) < 1:
            return None
        nu, q = self.map.integrate(_SCRATCH_FRAME, use_color, depthExtrinsic, depthCamera, chunksIntersecting,
                                   integrate_flag, needsUpdateFlag)
        if needsUpdateFlag is not None:
            needsUpdateFlag[:] = nu
        if keyframeID >= 0:  # Structure/Chisel.h:244-247
            for i in np.nonzero((q > 0) & (nu != 0))[0]:
                cid = tuple(int(v) for v in chunksIntersecting[i])
                self.chunkManager._obs.setdefault(cid, {})[keyframeID] = float(q[i])
        return nu

    def _mark(self, ids, upd):
        for cid in ids[upd != 0]:
            c = (int(cid[0]), int(cid[1]), int(cid[2]))
            for d in _NEIGHBOURS:
                self.meshesToUpdate[(c[0] + d[0], c[1] + d[1], c[2] + d[2])] = True

    # Structure/Chisel.h:184-216
    def FinalizeIntegrateChunks(self, chunksIntersecting, needsUpdateFlag, newChunkFlag):
        """Returns validChunks.  Marks meshesToUpdate and garbage-collects."""
        ids = np.asarray(chunksIntersecting, np.int32).reshape(-1, 3)
        nu = np.asarray(needsUpdateFlag) != 0
        new = np.asarray(newChunkFlag) != 0
        self._mark(ids, nu)
        garbage = ids[(~nu) & new]
        self.GarbageCollect(garbage)
        return ids[nu].copy()

    # Structure/Chisel.h:472-477
    def GarbageCollect(self, chunks):
        chunks = np.asarray(chunks, np.int32).reshape(-1, 3)
        if len(chunks):
            self.map.remove_chunks(chunks)
        for cid in chunks:
            c = (int(cid[0]), int(cid[1]), int(cid[2]))
            self.meshesToUpdate.pop(c, None)
            self.chunkManager._obs.pop(c, None)

    # Structure/Chisel.cpp:191-196
    def UpdateAtlas(self, chunksToUpdate):
        self.atlas.UpdateBuffers(chunksToUpdate)
