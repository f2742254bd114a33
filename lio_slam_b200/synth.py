"""Seeded synthetic inputs for the scan-to-map path.

A closed planar world (ground, tall perimeter walls, axis-aligned boxes, slanted rectangles) is ray
cast by a spinning LiDAR model (16/32/64/128 beams x 1800 azimuth steps, column-major firing order
like a Velodyne, per-point relative `time`, ring id, range noise).  Everything is numpy and fully
determined by the seed, so the CPU oracle and the CUDA library see identical bytes.

Layouts follow the reference: scans are PointXYZIRT records (imageProjection.cpp:4-15, 32 B:
x y z pad intensity ring(u16) time pad) and clouds are pcl::PointXYZI records (utility.h:65, 32 B)
or packed float4 (x, y, z, intensity).
"""
from __future__ import annotations

import dataclasses
import numpy as np

BEAM_FOV_DEG = {16: 15.0, 32: 15.0, 64: 12.5, 128: 22.5}
HALF = 50.0        # world half extent (m)
WALL_H = 45.0      # perimeter wall height (m): every ray of every beam hits something < 100 m
SENSOR_Z = 1.8

XYZIRT_DTYPE = np.dtype(
    {"names": ["x", "y", "z", "intensity", "ring", "time"],
     "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
     "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
XYZI_DTYPE = np.dtype(
    {"names": ["x", "y", "z", "intensity"], "formats": ["<f4"] * 4,
     "offsets": [0, 4, 8, 16], "itemsize": 32})


def rpy_to_R(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """Rz(yaw) @ Ry(pitch) @ Rx(roll) in float64 — the convention of pcl::getTransformation."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, sy * sr + cy * sp * cr],
                     [sy * cp, cy * cr + sy * sp * sr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


@dataclasses.dataclass
class World:
    boxes_lo: np.ndarray   # (B,3)
    boxes_hi: np.ndarray   # (B,3)
    rect_c: np.ndarray     # (S,3) centre of slanted rectangles
    rect_u: np.ndarray     # (S,3) unit in-plane axes
    rect_v: np.ndarray
    rect_n: np.ndarray     # (S,3) unit normal
    rect_hu: np.ndarray    # (S,) half extents
    rect_hv: np.ndarray


def make_world(seed: int = 1234, n_boxes: int = 40, n_rects: int = 10) -> World:
    rng = np.random.default_rng(seed)
    # boxes keep clear of the 14 m disc around the origin where the sensor drives
    centres = []
    while len(centres) < n_boxes:
        c = rng.uniform(-HALF + 6, HALF - 6, size=2)
        if np.hypot(*c) > 16.0:
            centres.append(c)
    centres = np.array(centres)
    size = rng.uniform(2.0, 8.0, size=(n_boxes, 3))
    lo = np.column_stack([centres - size[:, :2] / 2, np.zeros(n_boxes)])
    hi = np.column_stack([centres + size[:, :2] / 2, size[:, 2]])
    rc, ru, rv, rn, hu, hv = [], [], [], [], [], []
    while len(rc) < n_rects:
        c = np.append(rng.uniform(-HALF + 8, HALF - 8, size=2), rng.uniform(1.5, 5.0))
        if np.hypot(c[0], c[1]) < 16.0:
            continue
        yaw = rng.uniform(0, 2 * np.pi)
        tilt = rng.uniform(np.deg2rad(20), np.deg2rad(70))
        R = rpy_to_R(0.0, tilt, yaw)
        rc.append(c); ru.append(R[:, 0]); rv.append(R[:, 1]); rn.append(R[:, 2])
        hu.append(rng.uniform(1.5, 4.0)); hv.append(rng.uniform(1.5, 4.0))
    return World(lo, hi, np.array(rc), np.array(ru), np.array(rv), np.array(rn), np.array(hu), np.array(hv))


def _raycast(world: World, o: np.ndarray, d: np.ndarray) -> np.ndarray:
    """Nearest positive hit distance for rays o + t d (o: (3,), d: (N,3) unit). float64."""
    n = d.shape[0]
    t_best = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        # ground z = 0
        t = -o[2] / d[:, 2]
        ok = (t > 1e-6)
        px = o[0] + t * d[:, 0]; py = o[1] + t * d[:, 1]
        ok &= (np.abs(px) <= HALF) & (np.abs(py) <= HALF)
        t_best = np.where(ok & (t < t_best), t, t_best)
        # perimeter walls
        for axis, sign in ((0, 1.0), (0, -1.0), (1, 1.0), (1, -1.0)):
            t = (sign * HALF - o[axis]) / d[:, axis]
            other = 1 - axis
            po = o[other] + t * d[:, other]
            pz = o[2] + t * d[:, 2]
            ok = (t > 1e-6) & (np.abs(po) <= HALF) & (pz >= 0) & (pz <= WALL_H)
            t_best = np.where(ok & (t < t_best), t, t_best)
        # boxes (slab test), chunked over rays to bound memory
        inv = 1.0 / d
        for s in range(0, n, 65536):
            e = min(n, s + 65536)
            t1 = (world.boxes_lo[None, :, :] - o[None, None, :]) * inv[s:e, None, :]
            t2 = (world.boxes_hi[None, :, :] - o[None, None, :]) * inv[s:e, None, :]
            tmin = np.nanmax(np.minimum(t1, t2), axis=2)
            tmax = np.nanmin(np.maximum(t1, t2), axis=2)
            hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
            tb = np.where(hit, tmin, np.inf).min(axis=1)
            t_best[s:e] = np.minimum(t_best[s:e], tb)
        # slanted rectangles
        for k in range(world.rect_c.shape[0]):
            denom = d @ world.rect_n[k]
            t = ((world.rect_c[k] - o) @ world.rect_n[k]) / denom
            p = o[None, :] + t[:, None] * d - world.rect_c[k][None, :]
            ok = (t > 1e-6) & (np.abs(p @ world.rect_u[k]) <= world.rect_hu[k]) & \
                 (np.abs(p @ world.rect_v[k]) <= world.rect_hv[k])
            t_best = np.where(ok & (t < t_best), t, t_best)
    return t_best


def beam_directions(beams: int, cols: int = 1800) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Unit directions in the sensor frame, column-major firing order: i = col*beams + ring."""
    fov = np.deg2rad(BEAM_FOV_DEG[beams])
    elev = np.linspace(-fov, fov, beams)
    az = -np.arange(cols) * (2 * np.pi / cols)          # clockwise like a Velodyne
    col = np.repeat(np.arange(cols), beams)
    ring = np.tile(np.arange(beams), cols)
    ce = np.cos(elev[ring]); se = np.sin(elev[ring])
    d = np.column_stack([ce * np.cos(az[col]), ce * np.sin(az[col]), se])
    return d, ring.astype(np.uint16), col


def make_scan(world: World, pose6, beams: int, seed: int, cols: int = 1800, max_range: float = 100.0,
              noise: float = 0.02) -> np.ndarray:
    """One LiDAR sweep taken at pose6 = (roll,pitch,yaw,x,y,z); returns XYZIRT records in the
    sensor frame (misses and returns beyond max_range dropped, as a driver would)."""
    rng = np.random.default_rng(seed)
    d, ring, col = beam_directions(beams, cols)
    R = rpy_to_R(*[float(v) for v in pose6[:3]])
    o = np.asarray(pose6[3:6], dtype=np.float64)
    t = _raycast(world, o, d @ R.T)
    t = t + rng.normal(0.0, noise, size=t.shape)
    keep = np.isfinite(t) & (t < max_range) & (t > 0.5)
    p = (d * t[:, None])[keep]
    out = np.zeros(int(keep.sum()), dtype=XYZIRT_DTYPE)
    out["x"] = p[:, 0].astype(np.float32); out["y"] = p[:, 1].astype(np.float32); out["z"] = p[:, 2].astype(np.float32)
    out["intensity"] = rng.uniform(0, 100, size=keep.shape)[keep].astype(np.float32)
    out["ring"] = ring[keep]
    out["time"] = (0.1 * col[keep] / cols).astype(np.float32)
    return out


def xyzirt_to_xyzi(scan: np.ndarray) -> np.ndarray:
    out = np.zeros(scan.shape[0], dtype=XYZI_DTYPE)
    for f in ("x", "y", "z", "intensity"):
        out[f] = scan[f]
    return out


def to_packed(cloud: np.ndarray) -> np.ndarray:
    """structured XYZI / XYZIRT records -> contiguous (n,4) float32 (x,y,z,intensity)."""
    return np.ascontiguousarray(np.column_stack([cloud["x"], cloud["y"], cloud["z"], cloud["intensity"]]).astype(np.float32))


def from_packed(p4: np.ndarray) -> np.ndarray:
    out = np.zeros(p4.shape[0], dtype=XYZI_DTYPE)
    out["x"], out["y"], out["z"], out["intensity"] = p4[:, 0], p4[:, 1], p4[:, 2], p4[:, 3]
    return out


def transform_packed(p4: np.ndarray, pose6) -> np.ndarray:
    """float64 rigid transform of a packed cloud by pose6 (generator use only — not the parity path)."""
    R = rpy_to_R(*[float(v) for v in pose6[:3]])
    out = p4.astype(np.float64).copy()
    out[:, :3] = out[:, :3] @ R.T + np.asarray(pose6[3:6], dtype=np.float64)
    return out.astype(np.float32)


def voxel_numpy(p4: np.ndarray, leaf: float) -> np.ndarray:
    """Plain numpy voxel centroid filter for GENERATING maps (float64 sums; not the parity path)."""
    ijk = np.floor(p4[:, :3].astype(np.float64) / leaf).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    out = np.zeros((uniq.shape[0], 4), dtype=np.float64)
    for c in range(4):
        out[:, c] = np.bincount(inv, weights=p4[:, c].astype(np.float64), minlength=uniq.shape[0]) / cnt
    return out.astype(np.float32)


def path_pose(s: float) -> np.ndarray:
    """Ground-truth pose at arc length s (m) along a gentle loop of radius 8 m around the origin."""
    r = 8.0
    th = s / r
    yaw = th + np.pi / 2
    return np.array([0.01 * np.sin(0.7 * th), 0.015 * np.cos(0.5 * th), yaw,
                     r * np.cos(th), r * np.sin(th), SENSOR_Z + 0.02 * np.sin(th)], dtype=np.float64)


def make_local_map(world: World, beams: int, n_map: int, leaf: float, seed: int, s0: float = 0.0,
                   spacing: float = 0.5, max_poses: int = 64, cols: int = 1800) -> np.ndarray:
    """Union of sweeps from poses along the path, in the map frame, voxelised with `leaf` and cut by a
    seeded random subset to exactly n_map points (order kept).  Returns packed (n_map,4) float32."""
    rng = np.random.default_rng(seed)
    clouds = []
    k = 0
    vox = np.zeros((0, 4), dtype=np.float32)
    while k < max_poses:
        pose = path_pose(s0 - spacing * k)
        sc = make_scan(world, pose, beams, seed * 1000 + k, cols=cols)
        clouds.append(transform_packed(to_packed(sc), pose))
        k += 1
        if k in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64):
            vox = voxel_numpy(np.concatenate(clouds), leaf)
            if vox.shape[0] >= n_map:
                break
    if vox.shape[0] < n_map:
        raise RuntimeError(f"only {vox.shape[0]} map points from {k} poses; lower leaf or raise max_poses")
    sel = np.sort(rng.choice(vox.shape[0], size=n_map, replace=False))
    return np.ascontiguousarray(vox[sel])


def perturbed_guess(pose_gt, seed: int, rot_deg=(0.5, 0.5, 1.0), trans=(0.10, 0.10, 0.05)) -> np.ndarray:
    """Initial guess = ground truth perturbed by up to (rot_deg, trans), seeded."""
    rng = np.random.default_rng(seed)
    d = np.concatenate([np.deg2rad(rot_deg), trans]) * rng.uniform(-1, 1, size=6)
    return (np.asarray(pose_gt, dtype=np.float64) + d).astype(np.float32)


def make_imu_table(time_scan_cur: float, seed: int, rate_hz: float = 200.0, span: float = 0.1):
    """What imuDeskewInfo (imageProjection.cpp:359-418) leaves in imuTime / imuRotX/Y/Z for a sweep:
    Euler-integrated gyro (yaw rate 0.6 rad/s + 0.2 sin(2*pi*2t) on roll/pitch, noise 1e-3)."""
    rng = np.random.default_rng(seed)
    dt = 1.0 / rate_hz
    t = np.arange(time_scan_cur - 0.008, time_scan_cur + span + 0.012, dt)
    rel = t - time_scan_cur
    gx = 0.2 * np.sin(2 * np.pi * 2 * rel) + rng.normal(0, 1e-3, t.shape)
    gy = 0.2 * np.sin(2 * np.pi * 2 * rel + 1.0) + rng.normal(0, 1e-3, t.shape)
    gz = 0.6 + rng.normal(0, 1e-3, t.shape)
    rx = np.zeros_like(t); ry = np.zeros_like(t); rz = np.zeros_like(t)
    for k in range(1, t.shape[0]):
        h = t[k] - t[k - 1]
        rx[k] = rx[k - 1] + gx[k] * h
        ry[k] = ry[k - 1] + gy[k] * h
        rz[k] = rz[k - 1] + gz[k] * h
    return t, rx, ry, rz


def write_sequence(path: str, world: World, beams: int, n_scans: int, seed: int, cols: int = 600, step: float = 0.35,
                   guess_noise: float = 0.5):
    """A replay file for lio_slam_b200/host/replay_driver.cpp: n_scans sweeps along the path; the initial
    guess of each sweep is the ground truth perturbed like an IMU-odometry prediction (scan 0: exact).
    Returns the ground-truth poses (n_scans, 6)."""
    import struct
    gts = []
    with open(path, "wb") as f:
        f.write(struct.pack("<i", n_scans))
        for s in range(n_scans):
            gt = path_pose(step * s)
            gts.append(gt)
            scan = xyzirt_to_xyzi(make_scan(world, gt, beams, seed * 100 + s, cols=cols))
            scan_rec = np.zeros((scan.shape[0], 8), np.float32)
            scan_rec[:, 0], scan_rec[:, 1], scan_rec[:, 2] = scan["x"], scan["y"], scan["z"]
            scan_rec[:, 3] = 1.0
            scan_rec[:, 4] = scan["intensity"]
            guess = gt.astype(np.float32) if s == 0 else perturbed_guess(
                gt, seed * 7 + s, rot_deg=(0.2 * guess_noise, 0.2 * guess_noise, 0.5 * guess_noise),
                trans=(0.08 * guess_noise, 0.08 * guess_noise, 0.03 * guess_noise))
            f.write(struct.pack("<d", 0.1 * s))
            f.write(np.asarray(guess, np.float32).tobytes())
            f.write(struct.pack("<i", scan_rec.shape[0]))
            f.write(scan_rec.tobytes())
    return np.array(gts)
