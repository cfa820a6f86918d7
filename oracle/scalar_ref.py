"""Second, independent restatement of the fusion arithmetic in NumPy float32 scalars/rows.

TEST INFRASTRUCTURE ONLY.  Written from the behavioural spec rather
than from tf_oracle.cpp, and deliberately naive (one 8-voxel row at a time, np.float32 ops
that round after every operation).  tests/test_oracle_cpu.py checks the AVX2 oracle against
this on small cases; together with the hand-computed known-answer tests it is what stands
in for the golden vectors the reference does not ship.

Reference citations as in the oracle: ProjectionIntegrator.cpp:67-426 (voxel update),
Structure/ChunkManager.h:197-207,303-636 (bbox, culling), Structure/Chisel.cpp:52-110 (centroids),
QuadraticTruncator.h:45-48, ConstantWeighter.h:43-46, PinholeCamera.h:46-49.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
INT_MIN = -(2 ** 31)
SENTINEL = f32(-99999999999.0)


DOT3_LEFT_TO_RIGHT = False  # True: Eigen 3.2 association (see oracle/eigen_standin/Eigen/Core)


def dot3(a0, b0, a1, b1, a2, b2):
    """Eigen fixed-size 3-vector inner product, every op rounded to float: x0 + (x1 + x2) from
    Eigen 3.3 on, (x0 + x1) + x2 with Eigen 3.2."""
    p0, p1, p2 = f32(f32(a0) * f32(b0)), f32(f32(a1) * f32(b1)), f32(f32(a2) * f32(b2))
    if DOT3_LEFT_TO_RIGHT:
        return f32(f32(p0 + p1) + p2)
    return f32(p0 + f32(p1 + p2))


def rne_x86(x) -> int:
    """_mm256_cvtps_epi32 on one lane: round-half-even; NaN / out of range -> INT_MIN."""
    x = float(x)
    if math.isnan(x) or math.isinf(x) or abs(x) >= 2147483648.0:
        return INT_MIN
    r = int(np.rint(x))  # rint = round half to even
    return r


def trunc_dist(trunc, z) -> np.float32:
    q, l, c, s, _ = [f32(v) for v in trunc]
    v = float(q) * (float(f32(z)) ** 2) + float(f32(l * f32(z))) + float(c)
    return f32(abs(v) * float(s))


def intrinsics(cam):
    return f32(int(cam.fx)), f32(int(cam.fy)), f32(int(cam.cx)), f32(int(cam.cy))


def centroid(pose, x, y, z, res):
    """cen = (Rt * (x,y,z)) * res + res/2, component k uses column k of R."""
    R = np.asarray(pose, f32)[:3, :3]
    half = f32(f32(res) * f32(0.5))
    out = []
    for k in range(3):
        m = dot3(R[0, k], f32(x), R[1, k], f32(y), R[2, k], f32(z))
        out.append(f32(f32(m * f32(res)) + half))
    return out


class ScalarChunk:
    def __init__(self):
        self.sdf = np.full(512, 999.0, f32)
        self.weight = np.zeros(512, f32)
        self.color = np.zeros((512, 4), np.uint16)


def voxel_update(chunk: ScalarChunk, chunk_id, res, trunc, depth, rgba, quality, pose, cam, flag):
    """Returns (updated, qsum).  depth HxW f32, rgba HxWx4 u8 or None, quality HxW f32 or None."""
    res = f32(res)
    pose = np.asarray(pose, f32)
    R, t = pose[:3, :3], pose[:3, 3]
    fx, fy, cx, cy = intrinsics(cam)
    W, H = cam.width, cam.height
    cxh, cyh = f32(float(cx) + 0.5), f32(float(cy) + 0.5)
    origin = [f32(f32(8 * int(c)) * res) for c in chunk_id]
    e = [f32(origin[k] - t[k]) for k in range(3)]
    o = [dot3(R[0, k], e[0], R[1, k], e[1], R[2, k], e[2]) for k in range(3)]
    tr = trunc_dist(trunc, o[2])
    wd = f32(f32(trunc[4]) / f32(f32(2.0) * tr))
    if not flag:
        wd = f32(-wd)
    diag = f32(math.sqrt(3.0) * float(res))
    thr_c = f32(float(f32(diag / f32(2.0))) + 0.01)
    thr_p = f32(tr + diag)
    near, far = f32(cam.near), f32(cam.far)
    dflat = depth.reshape(-1)
    updated, qsum = False, f32(0.0)
    for p in range(64):
        y, z = p & 7, p >> 3
        u, v, cz, valid = [], [], [], []
        for x in range(8):
            cen = centroid(pose, x, y, z, res)
            c = [f32(o[k] + cen[k]) for k in range(3)]
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                pu = f32(f32(f32(c[0] / c[2]) * fx) + cxh)
                pv = f32(f32(f32(c[1] / c[2]) * fy) + cyh)
            ui, vi = rne_x86(pu), rne_x86(pv)
            u.append(ui), v.append(vi), cz.append(c[2])
            valid.append(0 < ui < W - 1 and 0 < vi < H - 1)
        if not any(valid):
            break  # `continue` before `pos++`: the chunk is finished for this frame
        d = [f32(dflat[v[i] * W + u[i]]) if valid[i] else f32(0.0) for i in range(8)]
        sd = [f32(d[i] - cz[i]) for i in range(8)]
        if rgba is not None:
            upd = [valid[i] and (sd[i] > -thr_c) and (thr_c > sd[i]) for i in range(8)]
            if any(u[i] < 0 or u[i] > W - 1 or v[i] < 0 or v[i] > H - 1 for i in range(8)):
                qsum = SENTINEL
            if any(upd):
                if quality is not None:
                    s = f32(0.0)
                    for i in range(8):
                        s = f32(s + (f32(quality.reshape(-1)[v[i] * W + u[i]]) if upd[i] else f32(0.0)))
                    qsum = f32(qsum + s)
                for i in range(8):
                    px = rgba.reshape(-1, 4)[v[i] * W + u[i]].astype(np.int64) if upd[i] else np.zeros(4, np.int64)
                    col = chunk.color[p * 8 + i].astype(np.int64)
                    if flag:
                        col = (col + px) & 0xFFFF
                        n16 = int(col[3]) - 65536 if col[3] >= 32768 else int(col[3])
                        if n16 > 120:
                            col = col >> 2
                    else:
                        col = (col - px) & 0xFFFF
                    chunk.color[p * 8 + i] = col.astype(np.uint16)
        inb = [(d[i] > near) and (far > d[i]) and (sd[i] > f32(-0.03)) and (thr_p > sd[i]) for i in range(8)]
        if any(inb):
            updated = True
            for i in range(8):
                idx = p * 8 + i
                nw = wd if inb[i] else f32(0.0)
                w0, s0 = chunk.weight[idx], chunk.sdf[idx]
                num = f32(f32(s0 * w0) + f32(sd[i] * nw))
                den = f32(f32(w0 + nw) + f32(1e-4))
                with np.errstate(divide="ignore", invalid="ignore"):
                    ns = f32(num / den)
                nwt = f32(w0 + nw)
                if nwt > f32(0.5):
                    chunk.sdf[idx], chunk.weight[idx] = ns, nwt
                else:
                    chunk.sdf[idx], chunk.weight[idx] = f32(999.0), f32(0.0)
    return updated, qsum


def boundary_ids(depth, pose, cam, res):
    """Vectorised over pixels; float32 ops in the reference's left-to-right order."""
    pose = np.asarray(pose, f32)
    R, t = pose[:3, :3], pose[:3, 3]
    fx, fy, cx, cy = intrinsics(cam)
    H, W = depth.shape
    dz = (depth.astype(f32) + f32(0.2)).astype(f32)
    jj = np.arange(W, dtype=f32)[None, :]
    ii = np.arange(H, dtype=f32)[:, None]
    X = (((jj - cx) / fx).astype(f32) * dz).astype(f32)
    Y = (((ii - cy) / fy).astype(f32) * dz).astype(f32)
    inv = f32(f32(1.0) / f32(f32(8) * f32(res)))
    lo, hi = [], []
    for k in range(3):
        v = ((((R[k, 0] * X).astype(f32) + (R[k, 1] * Y).astype(f32)).astype(f32) + (R[k, 2] * dz).astype(f32)).astype(f32)
             + t[k]).astype(f32)
        lo.append(int(math.floor(float(f32(min(f32(1e8), v.min()) * inv)))))
        hi.append(int(math.floor(float(f32(max(f32(-1e8), v.max()) * inv)))))
    return np.array(lo, np.int32), np.array(hi, np.int32)


def corner_test(o, depth, cam, dtp, dtn, offs):
    fx, fy, cx, cy = intrinsics(cam)
    W, H = cam.width, cam.height
    if not (o[2] > f32(cam.near) and f32(cam.far) > o[2]):
        return False
    dflat = depth.reshape(-1)
    for off in offs:
        c = [f32(o[k] + off[k]) for k in range(3)]
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            u = rne_x86(f32(f32(f32(c[0] / c[2]) * fx) + cx))
            v = rne_x86(f32(f32(f32(c[1] / c[2]) * fy) + cy))
        if 1 < u < W - 1 and 1 < v < H - 1:
            sd = f32(f32(dflat[v * W + u]) - c[2])
            if sd > f32(-dtn) and f32(dtp) > sd:
                return True
    return False


def observed_ids(depth, pose, cam, res, trunc):
    """GetChunkIDsObservedByCamera, scalar (slow: use small images / coarse voxels)."""
    res = f32(res)
    pose = np.asarray(pose, f32)
    R, t = pose[:3, :3], pose[:3, 3]
    lo, hi = boundary_ids(depth, pose, cam, res)
    diag, step, neg = f32(f32(f32(8) * res) / f32(2)), 4, f32(0.03)
    if float(res) > 0.01:
        diag = f32(float(f32(f32(8) * res)) * math.sqrt(3.0))
        step = 1
        neg = f32(0.05 * float(res) / 0.005)
    tau = [dot3(R[0, k], t[0], R[1, k], t[1], R[2, k], t[2]) for k in range(3)]
    r = [[f32(f32(R[k, i] * f32(8.0)) * res) for i in range(3)] for k in range(3)]  # r[k][i] = Rt(i,k)*8*res
    half = f32(res * f32(0.5))
    off_c, off_f = [], []
    for x in range(2):
        for y in range(2):
            for z in range(2):
                rc = [dot3(R[0, k], f32(8 * x), R[1, k], f32(8 * y), R[2, k], f32(8 * z)) for k in range(3)]
                off_c.append([f32(f32(f32(rc[k] * res) * f32(step)) + half) for k in range(3)])
                off_f.append([f32(f32(f32(rc[k] * res) * f32(1.0)) + half) for k in range(3)])
    out = []
    for x in range(lo[0] - 1, hi[0] + 2, step):
        ox = [f32(f32(r[0][k] * f32(x)) - tau[k]) for k in range(3)]
        for y in range(lo[1] - 1, hi[1] + 2, step):
            oy = [f32(ox[k] + f32(r[1][k] * f32(y))) for k in range(3)]
            for z in range(lo[2] - 1, hi[2] + 2, step):
                o = [f32(oy[k] + f32(f32(z) * r[2][k])) for k in range(3)]
                tr = trunc_dist(trunc, o[2])
                if not corner_test(o, depth, cam, f32(tr + f32(diag * f32(step))), f32(neg + f32(diag * f32(step))), off_c):
                    continue
                for i in range(x, x + step):
                    for j in range(y, y + step):
                        for k2 in range(z, z + step):
                            g = [f32(f32(8 * i) * res), f32(f32(8 * j) * res), f32(f32(8 * k2) * res)]
                            oc = [f32(dot3(R[0, k], g[0], R[1, k], g[1], R[2, k], g[2]) - tau[k]) for k in range(3)]
                            tr2 = trunc_dist(trunc, oc[2])
                            if corner_test(oc, depth, cam, f32(tr2 + diag), f32(neg + diag), off_f):
                                out.append((i, j, k2))
    return np.array(out, np.int32).reshape(-1, 3)
