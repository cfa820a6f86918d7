"""Deterministic synthetic RGB-D sequences for the fusion hot path.

A 6x4x3 m box room with axis-aligned boxes is ray-cast analytically to z-depth images
(float32 metres, 0 = invalid, as Frame::refined_depth, GCSLAM/frame.h:35-66), procedural
RGB, an all-valid colour mask and an observation-quality plane, with known camera->world
poses.  Layout on disk (write_dataset) follows the reference's offline dataset:
calib.txt / associate.txt / depth PNG(16 bit) / rgb PNG (Tools/DatasetWrapper.hpp:55-162).

Everything is computed with torch ops so that the same code runs on the CPU (tests, golden
fixtures) and on the GPU (bench set-up); outputs are returned as NumPy arrays.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

# room interior [lo, hi] and furniture boxes (metres, z up)
ROOM = ((0.0, 0.0, 0.0), (6.0, 4.0, 3.0))
BOXES = (
    ((0.6, 0.5, 0.0), (1.6, 1.3, 0.9)),
    ((4.3, 0.4, 0.0), (5.5, 1.0, 1.8)),
    ((2.4, 3.1, 0.0), (3.8, 3.8, 0.75)),
    ((0.3, 2.6, 0.0), (0.9, 3.7, 2.0)),
    ((5.0, 2.5, 0.0), (5.8, 3.5, 0.5)),
    ((2.7, 0.2, 0.9), (3.3, 0.5, 1.5)),
)
# building-scale floor plan for C4 (corridor + rooms), metres
BUILDING = ((0.0, 0.0, 0.0), (40.0, 25.0, 3.0))


@dataclass
class Camera:
    """chisel::PinholeCamera inputs (SetIntrinsics/SetWidth/SetHeight/SetNearPlane/SetFarPlane,
    GCFusion/MobileFusion.h:253-257)."""

    width: int = 640
    height: int = 480
    fx: float = 525.0
    fy: float = 525.0
    cx: float = 319.5
    cy: float = 239.5
    near: float = 0.01
    far: float = 5.0

    def scaled(self, s: float) -> "Camera":
        return Camera(int(self.width * s), int(self.height * s), self.fx * s, self.fy * s,
                      (self.cx + 0.5) * s - 0.5, (self.cy + 0.5) * s - 0.5, self.near, self.far)


@dataclass
class Frame:
    index: int
    pose: np.ndarray            # 4x4 float32 camera->world (pose_sophus[0])
    depth: np.ndarray           # HxW float32
    rgb: np.ndarray | None      # HxWx3 uint8 (key-frames only)
    color_valid: np.ndarray | None  # HxW uint8
    quality: np.ndarray | None  # HxW float32
    is_keyframe: bool
    pose_old: np.ndarray | None = None  # pose_sophus[1] (drifted) for loop-closure configs

    def rgba(self) -> np.ndarray | None:
        """RGBA plane as ReIntegrateKeyframe packs it (GCFusion/MobileFusion.cpp:151-162)."""
        if self.rgb is None:
            return None
        h, w, _ = self.rgb.shape
        out = np.zeros((h, w, 4), np.uint8)
        valid = self.color_valid > 0 if self.color_valid is not None else np.ones((h, w), bool)
        out[..., :3] = np.where(valid[..., None], self.rgb, 0)
        out[..., 3] = valid
        return out


def orbit_pose(k: int, n: int, center=(3.0, 2.0, 1.5), radius=1.0, laps=1.0, bob=0.15, pitch=0.12):
    """Camera->world pose k of an n-pose outward-looking orbit (x right, y down, z forward)."""
    th = 2.0 * math.pi * laps * k / n
    pos = np.array([center[0] + radius * math.cos(th), center[1] + radius * math.sin(th),
                    center[2] + bob * math.sin(3.0 * th)], np.float64)
    ph = pitch * math.sin(2.0 * th + 0.7)
    fwd = np.array([math.cos(th) * math.cos(ph), math.sin(th) * math.cos(ph), math.sin(ph)])
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    T = np.eye(4, dtype=np.float64)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, pos
    return T.astype(np.float32)


def walk_pose(k: int, n: int):
    """Pose k of an n-pose walk through the C4 building (a rounded rectangle loop)."""
    u = (k / n) * 4.0
    seg, f = int(u) % 4, u - int(u)
    pts = [(5.0, 5.0), (35.0, 5.0), (35.0, 20.0), (5.0, 20.0)]
    a, b = np.array(pts[seg]), np.array(pts[(seg + 1) % 4])
    p = a + (b - a) * f
    d = (b - a) / np.linalg.norm(b - a)
    yaw = math.atan2(d[1], d[0]) + 0.6 * math.sin(2 * math.pi * f * 3)
    fwd = np.array([math.cos(yaw), math.sin(yaw), 0.05 * math.sin(2 * math.pi * f * 5)])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, (p[0], p[1], 1.5)
    return T.astype(np.float32)


def _slab(o, d, lo, hi):
    inv = 1.0 / d
    t1 = (lo - o) * inv
    t2 = (hi - o) * inv
    tmin = torch.minimum(t1, t2).amax(dim=-1)
    tmax = torch.maximum(t1, t2).amin(dim=-1)
    return tmin, tmax


def _building_boxes():
    """Interior walls/pillars for the C4 floor plan (deterministic)."""
    rng = np.random.RandomState(3)
    out = []
    for ix in range(8):
        for iy in range(5):
            x0, y0 = 2.0 + ix * 4.7, 1.5 + iy * 4.6
            w, h = rng.uniform(0.6, 2.0), rng.uniform(0.5, 1.8)
            hz = rng.uniform(0.6, 2.6)
            out.append(((x0, y0, 0.0), (x0 + w, y0 + h, hz)))
    return tuple(out)


_BUILDING_BOXES = None


def render(pose: np.ndarray, cam: Camera, *, color: bool, device="cpu", scene="room",
           noise_sigma: float = 0.0, noise_seed: int = 1, invalid_border: int | None = None):
    """Ray-cast one frame.  Returns depth (HxW f32) and, if color, rgb (HxWx3 u8) and
    quality (HxW f32).  A band of `invalid_border` pixels (default W/80, i.e. 8 px at 640)
    around the image has depth 0 like a real sensor's refined depth; the reference's
    depth bbox counts such pixels at 0.2 m (Structure/ChunkManager.h:325-351), which is what
    makes its candidate grid reach from the camera to the surfaces."""
    global _BUILDING_BOXES
    dev = torch.device(device)
    f64 = torch.float64
    H, W = cam.height, cam.width
    jj = torch.arange(W, dtype=f64, device=dev)
    ii = torch.arange(H, dtype=f64, device=dev)
    dx = ((jj - cam.cx) / cam.fx)[None, :].expand(H, W)
    dy = ((ii - cam.cy) / cam.fy)[:, None].expand(H, W)
    dc = torch.stack([dx, dy, torch.ones_like(dx)], dim=-1)  # z-depth parametrisation
    T = torch.as_tensor(np.asarray(pose, np.float64), device=dev)
    d = dc @ T[:3, :3].T
    d = torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    o = T[:3, 3]
    if scene == "room":
        room, boxes = ROOM, BOXES
    else:
        if _BUILDING_BOXES is None:
            _BUILDING_BOXES = _building_boxes()
        room, boxes = BUILDING, _BUILDING_BOXES
    lo = torch.tensor(room[0], dtype=f64, device=dev)
    hi = torch.tensor(room[1], dtype=f64, device=dev)
    _, t_room = _slab(o, d, lo, hi)
    t = t_room
    sid = torch.zeros((H, W), dtype=torch.int64, device=dev)
    for b, (blo, bhi) in enumerate(boxes):
        tmin, tmax = _slab(o, d, torch.tensor(blo, dtype=f64, device=dev),
                           torch.tensor(bhi, dtype=f64, device=dev))
        hit = (tmax >= tmin) & (tmin > 1e-6) & (tmin < t)
        t = torch.where(hit, tmin, t)
        sid = torch.where(hit, torch.full_like(sid, b + 1), sid)
    depth = t.clone()
    if noise_sigma > 0:
        g = torch.Generator().manual_seed(noise_seed)
        depth = depth + (torch.randn((H, W), generator=g, dtype=f64) * noise_sigma).to(dev)
    depth = torch.where((depth > cam.near) & (depth < cam.far), depth, torch.zeros_like(depth))
    b = max(1, W // 80) if invalid_border is None else invalid_border
    if b > 0:
        depth[:b, :] = 0
        depth[-b:, :] = 0
        depth[:, :b] = 0
        depth[:, -b:] = 0
    depth32 = depth.to(torch.float32)
    if not color:
        return depth32.cpu().numpy(), None, None
    p = o + d * t[..., None]
    # procedural texture: 25 cm checker modulated by a smooth gradient, per-surface hue
    cell = torch.floor(p * 4.0 + 1e-6).sum(-1).to(torch.int64)
    chk = (cell % 2).to(f64)
    hue = (sid.to(f64) * 0.61803398875) % 1.0
    base = torch.stack([0.55 + 0.4 * torch.cos(6.2831853 * (hue + s)) for s in (0.0, 0.33, 0.67)], -1)
    grad = 0.75 + 0.25 * torch.sin(p[..., 0:1] * 1.3 + p[..., 1:2] * 0.7 + p[..., 2:3] * 2.1)
    col = base * (0.55 + 0.45 * chk[..., None]) * grad
    rgb = (col.clamp(0, 1) * 255.0).round().to(torch.uint8)
    # quality surrogate: view-angle cosine times inverse depth (positive, smooth)
    n_d = d / d.norm(dim=-1, keepdim=True)
    cosv = (n_d * T[:3, 2]).sum(-1).abs()
    quality = (cosv / (1.0 + t)).to(torch.float32)
    quality = torch.where(depth32 > 0, quality, torch.zeros_like(quality))
    return depth32.cpu().numpy(), rgb.cpu().numpy(), quality.cpu().numpy()


def drift_pose(pose: np.ndarray, k: int, seed: int = 2, sigma_t=0.002, sigma_r_deg=0.05):
    """pose_sophus[1] for loop-closure configs: truth composed with a random-walk drift
    (translation sigma 2 mm/kf, rotation sigma 0.05 deg/kf), deterministic in (k, seed)."""
    rng = np.random.RandomState(seed)
    steps_t = rng.normal(0.0, sigma_t, size=(k + 1, 3)).sum(0)
    steps_r = np.deg2rad(rng.normal(0.0, sigma_r_deg, size=(k + 1, 3)).sum(0))
    ax, ay, az = steps_r
    Rx = np.array([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    Ry = np.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    Rz = np.array([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]])
    D = np.eye(4)
    D[:3, :3] = Rz @ Ry @ Rx
    D[:3, 3] = steps_t
    return (np.asarray(pose, np.float64) @ D).astype(np.float32)


@dataclass
class Sequence:
    cam: Camera
    frames: list = field(default_factory=list)
    keyframe_every: int = 10


def make_sequence(n_frames=300, *, cam: Camera | None = None, keyframe_every=10, device="cpu",
                  scene="room", total=None, noise_sigma=0.0, with_drift=False, start=0) -> Sequence:
    """C1/C2 sequence: orbit of `total` poses (default n_frames), every `keyframe_every`-th
    frame a key-frame with colour + quality, the others depth-only local frames."""
    cam = cam or Camera()
    total = total or n_frames
    seq = Sequence(cam, [], keyframe_every)
    for k in range(start, start + n_frames):
        pose = orbit_pose(k, total) if scene == "room" else walk_pose(k, total)
        is_kf = (k % keyframe_every) == 0
        depth, rgb, q = render(pose, cam, color=is_kf, device=device, scene=scene,
                               noise_sigma=noise_sigma, noise_seed=1 + k)
        fr = Frame(k, pose, depth, rgb, np.ones(depth.shape, np.uint8) if is_kf else None, q, is_kf)
        if with_drift:
            fr.pose_old = drift_pose(pose, k // keyframe_every)
        seq.frames.append(fr)
    return seq


def write_dataset(seq: Sequence, folder: str, depth_scale=1000.0):
    """Write the reference's offline dataset layout (Tools/DatasetWrapper.hpp:85-158):
    calib.txt `W H fx fy cx cy d0..d4 depth_scale max_depth`, associate.txt
    `tRGB rgb.png tDepth depth.png`, groundtruth.txt in TUM format (BasicAPI.cpp:82-87)."""
    import os

    import cv2

    os.makedirs(os.path.join(folder, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(folder, "depth"), exist_ok=True)
    c = seq.cam
    with open(os.path.join(folder, "calib.txt"), "w") as f:
        f.write(f"{c.width} {c.height} {c.fx} {c.fy} {c.cx} {c.cy} 0 0 0 0 0 {depth_scale} {c.far}\n")
    with open(os.path.join(folder, "associate.txt"), "w") as fa, \
            open(os.path.join(folder, "groundtruth.txt"), "w") as fg:
        for fr in seq.frames:
            ts = fr.index / 30.0
            cv2.imwrite(os.path.join(folder, f"depth/{fr.index:06d}.png"),
                        np.round(fr.depth * depth_scale).astype(np.uint16))
            rgb = fr.rgb if fr.rgb is not None else np.zeros((c.height, c.width, 3), np.uint8)
            cv2.imwrite(os.path.join(folder, f"rgb/{fr.index:06d}.png"), rgb[..., ::-1])
            fa.write(f"{ts:.6f} rgb/{fr.index:06d}.png {ts:.6f} depth/{fr.index:06d}.png\n")
            R, t = fr.pose[:3, :3].astype(np.float64), fr.pose[:3, 3]
            qw = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
            qx = (R[2, 1] - R[1, 2]) / (4 * qw)
            qy = (R[0, 2] - R[2, 0]) / (4 * qw)
            qz = (R[1, 0] - R[0, 1]) / (4 * qw)
            fg.write(f"{ts:.6f} {t[0]:.6f} {t[1]:.6f} {t[2]:.6f} {qx:.6f} {qy:.6f} {qz:.6f} {qw:.6f}\n")
