"""GPU parity of liogpu_make_scancontext against the oracle's restatement of SCManager::makeScancontext and its
ring / sector keys (include/Scancontext.cpp:151-225): every bin and key bit-equal."""
import numpy as np
import pytest

from lio_slam_b200 import synth

pytestmark = pytest.mark.gpu


def check(gpu, oracle, cloud, **kw):
    want = oracle.make_scancontext(cloud, **kw)
    got = gpu.make_scancontext(cloud, **kw)
    for a, b, name in zip(got, want, ("desc", "ringkey", "sectorkey")):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), name
    return got


@pytest.mark.parametrize("beams,seed", [(16, 1), (64, 2), (128, 3)])
def test_scancontext_of_sweeps(gpu, oracle, world, beams, seed):
    scan = synth.to_packed(synth.make_scan(world, synth.path_pose(0.7 * seed), beams, seed=50 + seed))
    desc, rk, sk = check(gpu, oracle, scan)
    assert (desc != 0).sum() > 200 and desc.max() > 2.0


def test_scancontext_bin_edges_and_degenerate_points(gpu, oracle):
    rng = np.random.default_rng(4)
    ang = np.deg2rad(np.arange(0, 360, 6.0))          # exactly on the sector boundaries
    r = np.arange(4.0, 84.0, 4.0)                     # exactly on the ring boundaries, incl. the last ring edge
    pts = [[rr * np.cos(a), rr * np.sin(a), 0.3 * k] for k, rr in enumerate(r) for a in ang]
    pts += [[0, 0, 1], [0, 5, 1], [5, 0, 1], [-5, 0, 1], [0, -5, 1], [80.0, 0, 3], [80.0001, 0, 9], [56.57, 56.57, 4],
            [1, 1, -1500.0], [2, 2, -1002.0], [3, 3, np.nan], [np.inf, 1, 1], [1e-30, -1e-30, 0.5]]
    cloud = np.c_[np.array(pts, np.float64), np.zeros(len(pts))].astype(np.float32)
    check(gpu, oracle, cloud)
    check(gpu, oracle, cloud, lidar_height=0.0, max_radius=40.0)
    big = np.c_[rng.uniform(-90, 90, (300000, 2)), rng.normal(0, 2, (300000, 1)), np.zeros((300000, 1))].astype(np.float32)
    check(gpu, oracle, big)
    desc, rk, sk = check(gpu, oracle, np.zeros((0, 4), np.float32))
    assert not desc.any() and not rk.any() and not sk.any()


def test_scancontext_of_resident_deskewed_sweep(oracle, world):
    """SCInputType::SINGLE_SCAN_FULL (mapOptmization.cpp:2151-2156): the descriptor of the deskewed sweep, read in
    place from HBM; the resident cloud stays usable afterwards"""
    from lio_slam_b200.liogpu import LioGpu, RESIDENT
    from oracle.oracle import DeskewParams
    g = LioGpu(n_scan=32, downsample_rate=1, point_filter_num=1, lidar_min_front=2.0, lidar_min_back=10.0,
               lidar_min_left=2.0, lidar_min_right=2.0, lidar_max_range=100.0, lidar_max_intensity=100.0)
    try:
        scan = synth.make_scan(world, synth.path_pose(0.4), 32, seed=77)
        t0 = 1700000000.0
        imu = synth.make_imu_table(t0, seed=5)
        dsk = oracle.deskew(scan, DeskewParams(32, 1, 1, 2.0, 10.0, 2.0, 2.0, 100.0, 100.0), t0, *imu, True)
        want = oracle.make_scancontext(dsk)
        g.deskew(scan, t0, *imu, True, keep_on_device=True)
        got = g.make_scancontext(RESIDENT)
        for a, b in zip(got, want):
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert g.resident_size() == dsk.shape[0]
        ds, _ = g.voxel_downsample(RESIDENT, 0.4)
        wds, _ = oracle.voxel_grid(dsk, 0.4)
        assert np.array_equal(ds.view(np.uint32), wds.view(np.uint32))
    finally:
        g.close()
