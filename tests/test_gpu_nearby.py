"""GPU parity of liogpu_extract_nearby against the oracle's restatement of extractNearby + extractCloud's guard
(mapOptmization.cpp:1519-1565): the ordered keyframe-index list must be identical."""
import numpy as np
import pytest

from lio_slam_b200 import synth

pytestmark = pytest.mark.gpu


def key_cloud(xyz):
    n = xyz.shape[0]
    return np.c_[xyz, np.arange(n)].astype(np.float32)


def check(gpu, oracle, key3d, t, t_cur, radius=50.0, density=2.0):
    want = oracle.extract_nearby(key3d, t, t_cur, radius, density)
    got, st = gpu.extract_nearby(key3d, t, t_cur, radius, density)
    assert st == 0
    assert np.array_equal(got, want), (got[:20], want[:20], got.shape, want.shape)
    return got


def test_nearby_on_a_loop(gpu, oracle):
    n = 400
    xyz = np.array([synth.path_pose(0.7 * k)[3:6] for k in range(n)])
    t = 100.0 + 0.35 * np.arange(n)
    ids = check(gpu, oracle, key_cloud(xyz), t, t[-1] + 0.1)
    assert ids[-1] == n - 29 and ids[-29] == n - 1          # 29 poses younger than 10 s, newest first
    check(gpu, oracle, key_cloud(xyz), t, t[-1] + 0.1, density=1.0)
    check(gpu, oracle, key_cloud(xyz), t, t[-1] + 50.0)     # nothing recent
    check(gpu, oracle, key_cloud(xyz), t, t[0] + 1.0)       # everything recent (time jumps back): all n appended
    check(gpu, oracle, key_cloud(xyz), t, t[-1] + 0.1, radius=5.0)


def test_nearby_long_drive_and_revisit(gpu, oracle):
    rng = np.random.default_rng(8)
    n = 3000
    step = np.c_[np.full(n, 0.9), rng.normal(0, 0.15, n), rng.normal(0, 0.01, n)]
    xyz = np.cumsum(step, axis=0)
    xyz[2000:] = xyz[2000:] - xyz[2000] + xyz[500] + [0.0, 0.3, 0.0]   # drive the same street again (loop)
    t = 10.0 + 0.5 * np.arange(n)
    ids = check(gpu, oracle, key_cloud(xyz), t, t[-1] + 0.2)
    assert (ids < 1600).any() and (ids > 2900).any()        # old keyframes of the first pass are selected too
    check(gpu, oracle, key_cloud(xyz[:700]), t[:700], t[699] + 0.2)    # straight drive: only the tail is in range
    # the recency rule reaches beyond the radius: those entries are dropped by extractCloud's guard
    check(gpu, oracle, key_cloud(xyz[:700]), t[:700], t[699] + 0.2, radius=4.0)


def test_nearby_ties_duplicates_and_tiny(gpu, oracle):
    # a robot standing still: identical key poses -> equal distances, equal voxels, nearest-pose ties
    xyz = np.zeros((50, 3))
    xyz[25:] = [1.0, 0.0, 0.0]
    t = np.arange(50) * 1.0
    check(gpu, oracle, key_cloud(xyz), t, 49.5)
    lattice = np.array([[i, j, 0.0] for i in range(-10, 11) for j in range(-10, 11)], float) * 2.0
    check(gpu, oracle, key_cloud(lattice), np.arange(lattice.shape[0]) * 0.3, lattice.shape[0] * 0.3)
    check(gpu, oracle, key_cloud(np.array([[3.0, 4.0, 0.5]])), np.array([5.0]), 5.1)          # first keyframe
    check(gpu, oracle, key_cloud(np.array([[0.0, 0, 0], [60.0, 0, 0]])), np.array([0.0, 1.0]), 1.5)


def test_nearby_ten_thousand_key_poses_and_pcl_layout(gpu, oracle):
    from lio_slam_b200.liogpu import W_NO_KEYFRAMES
    rng = np.random.default_rng(12)
    n = 10000
    ang = np.cumsum(rng.normal(0, 0.05, n))
    xyz = np.cumsum(np.c_[np.cos(ang), np.sin(ang), np.zeros(n)] * 0.8, axis=0)
    t = 0.4 * np.arange(n)
    key3d = key_cloud(xyz)
    want = check(gpu, oracle, key3d, t, t[-1] + 0.05)
    rec = np.zeros((n, 8), np.float32)      # pcl::PointXYZI records (stride 32, intensity at byte 16)
    rec[:, :3] = key3d[:, :3]
    rec[:, 3] = 1.0
    rec[:, 4] = key3d[:, 3]
    got, _ = gpu.extract_nearby(rec, t, t[-1] + 0.05)
    assert np.array_equal(got, want)
    got, st = gpu.extract_nearby(np.zeros((0, 4), np.float32), np.zeros(0), 0.0)
    assert st == W_NO_KEYFRAMES and got.shape[0] == 0


def test_nearby_feeds_build_local_map(gpu, oracle, world):
    """the id list drives extractCloud exactly as in the reference: same local map as the oracle's"""
    clouds, poses = [], []
    for k in range(12):
        p = synth.path_pose(1.5 * k)
        ds, _ = oracle.voxel_grid(synth.to_packed(synth.make_scan(world, p, 16, seed=900 + k, cols=300)), 0.4)
        clouds.append(ds)
        poses.append(p.astype(np.float32))
    poses = np.array(poses, np.float32)
    key3d = key_cloud(poses[:, 3:6].astype(np.float64))
    t = 2.0 * np.arange(12)
    ids, _ = gpu.extract_nearby(key3d, t, t[-1] + 0.1)
    assert np.array_equal(ids, oracle.extract_nearby(key3d, t, t[-1] + 0.1))
    gpu.keyframe_clear()
    for k, c in enumerate(clouds):
        gpu.keyframe_put(k, c)
    got, _ = gpu.build_local_map(ids, poses[ids], 0.5)
    want, _ = oracle.build_local_map([clouds[i] for i in ids], poses[ids], 0.5, threads=4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
