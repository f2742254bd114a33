"""Host-side sharding of the two configurations that shard naturally.  No collective on
the data path: every rank works on its own points and results are concatenated in rank order.

* batch offline mapping (BASELINE configs[4]): independent sequences -> `assign_sequences`.
* full-map VoxelGrid rebuild (configs[3]): points are split into contiguous ranges of the voxel index
  (iz, iy) — spatial slabs/tiles in PCL's output order — by a STABLE partition, each rank voxelises its
  range, and the rank outputs concatenated in rank order reproduce the single-GPU output bit for bit,
  because a voxel never spans two ranges and the within-voxel summation order is preserved.  The product
  makes this plan ON THE DEVICE (liogpu_voxel_tile, csrc/tile.cu); `plan_voxel_tiles` restates the same plan
  with numpy so that the CPU suite (oracle workers, gloo world_size 2) exercises the N > 1 logic and a GPU test
  can hold the device plan against it.
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


TILE_MAX_BINS = 65536


def plan_voxel_tiles(cloud4: np.ndarray, leaf: float, n_tiles: int):
    """Host restatement of the DEVICE-side tile plan of liogpu_voxel_tile (csrc/tile.cu), for the CPU tests of the
    N > 1 path (the product plans on the GPU; nothing here is on a measured path): coarse histogram of the voxel-row
    index (iz, iy) — at most 65,536 bins — and tile t = the bins between the first bin whose cumulative count reaches
    t*n/N and the one that reaches (t+1)*n/N: contiguous ranges of voxel rows balanced by point count.

    Returns (tile_of_point int32[n], bounds): tile t holds the bins [bounds[t], bounds[t+1]).
    Non-finite points are dropped by VoxelGrid anyway; they are sent to tile 0."""
    inv = np.float32(1.0) / np.float32(leaf)
    finite = np.isfinite(cloud4[:, :3]).all(axis=1)
    n_valid = int(finite.sum())
    if n_valid == 0:
        return np.zeros(cloud4.shape[0], np.int32), [0] * n_tiles + [1]
    y = cloud4[:, 1].astype(np.float32); z = cloud4[:, 2].astype(np.float32)
    iy = np.floor(y * inv).astype(np.int64); iz = np.floor(z * inv).astype(np.int64)
    iy[~finite] = 0; iz[~finite] = 0
    iy0 = int(np.floor(y[finite].min() * inv)); iz0 = int(np.floor(z[finite].min() * inv))
    dy = int(np.floor(y[finite].max() * inv)) - iy0 + 1
    dz = int(np.floor(z[finite].max() * inv)) - iz0 + 1
    nrows = dy * dz
    shift = 0
    while (((nrows - 1) >> shift) + 1) > TILE_MAX_BINS:
        shift += 1
    nbins = ((nrows - 1) >> shift) + 1
    bins = ((iz - iz0) * dy + (iy - iy0)) >> shift
    hist = np.bincount(bins[finite], minlength=nbins)
    cum = np.concatenate([[0], np.cumsum(hist)])[:-1]          # exclusive prefix
    bounds = [0]
    for t in range(1, n_tiles):
        want = (n_valid * t) // n_tiles
        b = int(np.searchsorted(cum, want, side="left"))       # first bin with cum[bin] >= want
        bounds.append(max(min(b, nbins), bounds[-1]))
    bounds.append(nbins)
    tile = (np.searchsorted(np.array(bounds[1:-1], dtype=np.int64), bins, side="right")).astype(np.int32)
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
