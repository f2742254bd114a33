"""The synthetic LiDAR model of synth.py (same closed world, same beam table, same record layout) evaluated with torch
on the GPU, for the workloads that need thousands of sweeps (BASELINE configs[4]: 8 sequences x 1000 64-beam sweeps —
the numpy ray caster takes ~0.2 s per sweep).  Bench / test INPUT generation only: nothing here is on the measured
path, and the CUDA library and the CPU oracle are always fed the same bytes (the generated records are copied to
the host once and handed to both).  Noise comes from a seeded torch generator, so a (seed, sweep) pair always gives
the same records on the same software stack; it is not bit-identical to synth.make_scan's numpy noise."""
from __future__ import annotations

import numpy as np
import torch

from . import synth


class TorchWorld:
    def __init__(self, world: synth.World, device):
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64, device=device)
        self.device = device
        self.boxes_lo, self.boxes_hi = f(world.boxes_lo), f(world.boxes_hi)
        self.rect_c, self.rect_u, self.rect_v, self.rect_n = f(world.rect_c), f(world.rect_u), f(world.rect_v), f(world.rect_n)
        self.rect_hu, self.rect_hv = f(world.rect_hu), f(world.rect_hv)
        self._beams = {}

    def beams(self, beams: int, cols: int):
        key = (beams, cols)
        if key not in self._beams:
            d, ring, col = synth.beam_directions(beams, cols)
            self._beams[key] = (torch.as_tensor(d, dtype=torch.float64, device=self.device),
                                torch.as_tensor(ring.astype(np.int32), device=self.device),
                                torch.as_tensor((0.1 * col / cols).astype(np.float32), device=self.device))
        return self._beams[key]


def _raycast(w: TorchWorld, o: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """Nearest positive hit distance of the rays o + t d (the arithmetic of synth._raycast, float64)."""
    inf = torch.full((d.shape[0],), float("inf"), dtype=torch.float64, device=d.device)
    H = synth.HALF
    t = -o[2] / d[:, 2]
    px = o[0] + t * d[:, 0]; py = o[1] + t * d[:, 1]
    ok = (t > 1e-6) & (px.abs() <= H) & (py.abs() <= H)
    best = torch.where(ok, t, inf)
    for axis, sign in ((0, 1.0), (0, -1.0), (1, 1.0), (1, -1.0)):
        t = (sign * H - o[axis]) / d[:, axis]
        other = 1 - axis
        po = o[other] + t * d[:, other]
        pz = o[2] + t * d[:, 2]
        ok = (t > 1e-6) & (po.abs() <= H) & (pz >= 0) & (pz <= synth.WALL_H)
        best = torch.minimum(best, torch.where(ok, t, inf))
    inv = 1.0 / d
    t1 = (w.boxes_lo[None, :, :] - o[None, None, :]) * inv[:, None, :]
    t2 = (w.boxes_hi[None, :, :] - o[None, None, :]) * inv[:, None, :]
    tmin = torch.nan_to_num(torch.minimum(t1, t2), nan=-float("inf")).amax(dim=2)
    tmax = torch.nan_to_num(torch.maximum(t1, t2), nan=float("inf")).amin(dim=2)
    hit = (tmax >= tmin.clamp_min(0.0)) & (tmin > 1e-6)
    best = torch.minimum(best, torch.where(hit, tmin, inf[:, None].expand_as(tmin)).amin(dim=1))
    denom = d @ w.rect_n.T                                    # (N,S)
    t = ((w.rect_c - o[None, :]) * w.rect_n).sum(dim=1)[None, :] / denom
    p = o[None, None, :] + t[:, :, None] * d[:, None, :] - w.rect_c[None, :, :]
    ok = (t > 1e-6) & ((p * w.rect_u[None]).sum(dim=2).abs() <= w.rect_hu[None]) & \
         ((p * w.rect_v[None]).sum(dim=2).abs() <= w.rect_hv[None])
    best = torch.minimum(best, torch.where(ok, t, inf[:, None].expand_as(t)).amin(dim=1))
    return best


def make_scan_records(w: TorchWorld, pose6, beams: int, seed: int, cols: int = 1800, max_range: float = 100.0,
                      noise: float = 0.02) -> torch.Tensor:
    """One sweep taken at pose6 -> (n, 8) float32 tensor on the device whose bytes are PointXYZIRT records
    (x y z pad intensity ring|pad time pad), misses and returns beyond max_range dropped, firing order kept."""
    d, ring, tcol = w.beams(beams, cols)
    R = torch.as_tensor(synth.rpy_to_R(*[float(v) for v in pose6[:3]]), dtype=torch.float64, device=w.device)
    o = torch.as_tensor(np.asarray(pose6[3:6], dtype=np.float64), device=w.device)
    t = _raycast(w, o, d @ R.T)
    gen = torch.Generator(device=w.device)
    gen.manual_seed(int(seed))
    t = t + noise * torch.randn(t.shape, generator=gen, dtype=torch.float64, device=w.device)
    inten = 100.0 * torch.rand(t.shape, generator=gen, dtype=torch.float32, device=w.device)
    keep = torch.isfinite(t) & (t < max_range) & (t > 0.5)
    p = (d * t[:, None]).to(torch.float32)[keep]
    rec = torch.zeros((p.shape[0], 8), dtype=torch.float32, device=w.device)
    rec[:, 0:3] = p
    rec[:, 4] = inten[keep]
    rec[:, 5] = ring[keep].view(torch.float32)      # little endian: the u16 ring occupies bytes 20-21, 22-23 stay 0
    rec[:, 6] = tcol[keep]
    return rec


def records_to_numpy(rec: torch.Tensor) -> np.ndarray:
    """device records -> structured numpy array (synth.XYZIRT_DTYPE)"""
    return rec.cpu().numpy().view(np.uint8).reshape(-1, 32).view(synth.XYZIRT_DTYPE).reshape(-1)


def make_sequence(world: synth.World, beams: int, n_scans: int, seed: int, device, cols: int = 1800, step: float = 0.35,
                  s0: float = 0.0, guess_noise: float = 0.5, imu_yaw_rate: float = 0.02):
    """One drive along synth.path_pose: raw sweeps as ONE pinned host tensor of PointXYZIRT records plus per-sweep
    offsets, header stamps, initial guesses (ground truth perturbed like an IMU-odometry prediction; sweep 0 exact) and
    the IMU rotation tables imuDeskewInfo would leave (200 Hz, small rates: the sweeps are rendered from a static pose).
    -> dict(raw (pinned uint8 [total*32]), offs (n+1), times, guesses (n,6) f32, imu list of (4,k) f64, gts (n,6))"""
    w = TorchWorld(world, device)
    recs, offs, gts, guesses, times, imus = [], [0], [], [], [], []
    for s in range(n_scans):
        gt = synth.path_pose(s0 + step * s)
        r = make_scan_records(w, gt, beams, seed * 100003 + s, cols=cols)
        recs.append(r)
        offs.append(offs[-1] + int(r.shape[0]))
        gts.append(gt)
        guesses.append(gt.astype(np.float32) if s == 0 else synth.perturbed_guess(
            gt, seed * 7 + s, rot_deg=(0.2 * guess_noise, 0.2 * guess_noise, 0.5 * guess_noise),
            trans=(0.08 * guess_noise, 0.08 * guess_noise, 0.03 * guess_noise)))
        t = 0.1 * s
        times.append(t)
        rng = np.random.default_rng(seed * 31 + s)
        tt = np.arange(t - 0.008, t + 0.112, 0.005)
        gz = imu_yaw_rate + rng.normal(0, 1e-3, tt.shape)
        gx = rng.normal(0, 1e-3, tt.shape); gy = rng.normal(0, 1e-3, tt.shape)
        h = np.diff(tt, prepend=tt[0])
        imus.append(np.ascontiguousarray(np.stack([tt, np.cumsum(gx * h), np.cumsum(gy * h), np.cumsum(gz * h)])))
    total = offs[-1]
    raw = torch.empty((total, 8), dtype=torch.float32).pin_memory()
    torch.cat(recs, out=None).to("cpu", non_blocking=False) if False else None
    pos = 0
    for r in recs:
        raw[pos:pos + r.shape[0]].copy_(r)
        pos += r.shape[0]
    torch.cuda.synchronize(device)
    return dict(raw=raw, offs=np.array(offs, np.int64), times=np.array(times), guesses=np.array(guesses, np.float32),
                imu=imus, gts=np.array(gts))
