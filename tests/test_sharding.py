"""Multi-GPU host logic on CPU: spatial tiling of the full-map VoxelGrid rebuild (BASELINE configs[3]) and
the sequence-per-rank assignment (configs[4]); a world_size-2 gloo run exercises the N>1 path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lio_slam_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def biteq(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def keyframe_union(world, k=6, cols=300):
    clouds = []
    for i in range(k):
        p = synth.path_pose(-0.8 * i)
        clouds.append(synth.transform_packed(synth.to_packed(synth.make_scan(world, p, 32, seed=300 + i, cols=cols)), p))
    return np.concatenate(clouds)


@pytest.mark.parametrize("tiles", [1, 2, 3, 4, 8])
def test_tiled_voxel_equals_single(oracle, world, tiles):
    cloud = keyframe_union(world)
    want, ov = oracle.voxel_grid(cloud, 0.5)
    got, ov2 = sharding.voxel_downsample_sharded(cloud, 0.5, tiles, oracle.voxel_grid)
    assert not ov and not ov2
    assert biteq(got, want)
    tile, bounds = sharding.plan_voxel_tiles(cloud, 0.5, tiles)
    counts = np.bincount(tile, minlength=tiles)
    assert counts.sum() == cloud.shape[0] and (counts > 0).all()
    assert counts.max() <= 2.0 * cloud.shape[0] / tiles + 1000          # balanced by point count
    assert all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))


def test_tiled_voxel_overflow_guard(oracle, world):
    cloud = keyframe_union(world, k=2)
    want, ov = oracle.voxel_grid(cloud, 0.001)
    got, ov2 = sharding.voxel_downsample_sharded(cloud, 0.001, 4, oracle.voxel_grid)
    assert ov and ov2 and biteq(got, want) and biteq(got, cloud)


def test_sequence_assignment():
    for ws in (1, 2, 4, 8):
        seen = sorted(s for r in range(ws) for s in sharding.assign_sequences(8, ws, r))
        assert seen == list(range(8))
        assert max(len(sharding.assign_sequences(8, ws, r)) for r in range(ws)) == 8 // ws


def test_gloo_world_size_2():
    """Two CPU ranks (gloo): each voxelises its tiles with the oracle, rank 0 gathers in rank order."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r), WORLD_SIZE="2"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK" in outs[0], outs[0]
