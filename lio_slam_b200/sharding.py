"""Host-side sharding of the two configurations that shard naturally.  No collective on
the data path: every rank works on its own points and results are concatenated in rank order.

* batch offline mapping (BASELINE configs[4]): independent sequences -> `assign_sequences`.
* full-map VoxelGrid rebuild (configs[3]): points are split into contiguous ranges of the voxel index
  (iz, iy) — spatial slabs/tiles in PCL's output order — by a STABLE partition, each rank voxelises its
  range, and the rank outputs concatenated in rank order reproduce the single-GPU output bit for bit,
  because a voxel never spans two ranges and the within-voxel summation order is preserved.
"""
from __future__ import annotations

import numpy as np


def assign_sequences(n_sequences: int, world_size: int, rank: int) -> list[int]:
    """Sequence s runs on rank s mod world_size."""
    return [s for s in range(n_sequences) if s % world_size == rank]


def voxel_guard_fires(cloud4: np.ndarray, leaf: float) -> bool:
    """pcl::VoxelGrid's index-overflow guard on the WHOLE cloud, f32 arithmetic."""
    finite = np.isfinite(cloud4[:, :3]).all(axis=1)
    if not finite.any():
        return False
    p = cloud4[finite, :3].astype(np.float32)
    inv = np.float32(1.0) / np.float32(leaf)
    d = ((p.max(axis=0) - p.min(axis=0)) * inv).astype(np.int64) + 1
    return int(d[0]) * int(d[1]) * int(d[2]) > np.iinfo(np.int32).max


def plan_voxel_tiles(cloud4: np.ndarray, leaf: float, n_tiles: int):
    """Split the cloud into n_tiles contiguous ranges of the (iz, iy) voxel row index, balanced by point count.

    Returns (tile_of_point int32[n], bounds): tile t holds the voxel rows with key in [bounds[t], bounds[t+1]).
    Non-finite points are dropped by VoxelGrid anyway; they are sent to tile 0."""
    inv = np.float32(1.0) / np.float32(leaf)
    finite = np.isfinite(cloud4[:, :3]).all(axis=1)
    iy = np.floor(cloud4[:, 1].astype(np.float32) * inv).astype(np.int64)
    iz = np.floor(cloud4[:, 2].astype(np.float32) * inv).astype(np.int64)
    iy[~finite] = 0; iz[~finite] = 0
    iy0, iz0 = iy[finite].min(initial=0), iz[finite].min(initial=0)
    dy = int(iy[finite].max(initial=0) - iy0 + 1)
    key = (iz - iz0) * dy + (iy - iy0)
    order = np.sort(key[finite])
    bounds = [int(order[0]) if order.size else 0]
    for t in range(1, n_tiles):
        q = order[min(order.size - 1, (order.size * t) // n_tiles)] if order.size else 0
        bounds.append(max(int(q), bounds[-1]))
    bounds.append(int(order[-1]) + 1 if order.size else 1)
    tile = np.searchsorted(np.array(bounds[1:-1], dtype=np.int64), key, side="right").astype(np.int32)
    tile[~finite] = 0
    return tile, bounds


def shard_points(cloud4: np.ndarray, tile: np.ndarray, t: int) -> np.ndarray:
    """Stable selection: tile t's points in their original order (keeps the canonical summation order)."""
    return np.ascontiguousarray(cloud4[tile == t])


def voxel_downsample_sharded(cloud4: np.ndarray, leaf: float, n_tiles: int, voxel_fn):
    """Run voxel_fn(points, leaf) -> (out4, overflow) on every tile and concatenate in tile order.
    On the overflow guard of the whole cloud the input is returned unchanged, as PCL does (quirk q4)."""
    if voxel_guard_fires(cloud4, leaf):
        return cloud4.copy(), True
    tile, _ = plan_voxel_tiles(cloud4, leaf, n_tiles)
    outs = []
    for t in range(n_tiles):
        pts = shard_points(cloud4, tile, t)
        if pts.shape[0]:
            out, _ = voxel_fn(pts, leaf)
            outs.append(out)
    return (np.concatenate(outs) if outs else np.zeros((0, 4), np.float32)), False
