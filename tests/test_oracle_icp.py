"""CPU checks of the oracle's loop-closure ICP restatement (pcl::IterativeClosestPoint as configured at
mapOptmization.cpp:1111-1123): nearest-neighbour providers agree, a known rigid motion is
recovered, and one ICP step equals an independent numpy Kabsch/Umeyama step.  (PARITY UNPINNED: no PCL here.)"""
import numpy as np
import pytest

from lio_slam_b200 import synth


def rigid(rx, ry, rz, t):
    T = np.eye(4)
    T[:3, :3] = synth.rpy_to_R(rx, ry, rz)
    T[:3, 3] = t
    return T


def apply(T, c):
    out = c.copy()
    out[:, :3] = (c[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


@pytest.fixture(scope="module")
def clouds(world, oracle):
    p = synth.path_pose(2.0)
    tgt, _ = oracle.voxel_grid(synth.transform_packed(synth.to_packed(synth.make_scan(world, p, 32, seed=41, cols=600)), p), 0.4)
    return tgt


def test_oracle_icp_recovers_known_motion(oracle, clouds):
    tgt = clouds
    Tm = rigid(0.004, -0.003, 0.02, [0.35, -0.25, 0.05])
    src = apply(np.linalg.inv(Tm), tgt[::2])                   # moving src by Tm lands exactly on target points
    r = oracle.icp_align(src, tgt, threads=4)
    assert r["converged"] == 1 and r["state"] in (2, 3, 4) and r["iterations"] < 60
    assert np.abs(r["T"].astype(np.float64) - Tm).max() < 2e-3 and r["fitness_score"] < 1e-4
    assert r["n_correspondences"] == src.shape[0]


def test_oracle_icp_nn_providers_agree(oracle, clouds):
    from oracle.oracle import Oracle
    tgt = clouds
    src = apply(rigid(0.0, 0.0, 0.03, [0.6, 0.4, 0.0]), tgt[::5])
    a = oracle.icp_align(src, tgt, brute=True, threads=4)
    b = oracle.icp_align(src, tgt, brute=False, threads=4)
    c = oracle.icp_align(src, tgt, brute=False, threads=1)
    for other in (b, c):
        assert np.array_equal(a["T"], other["T"]) and a["iterations"] == other["iterations"]
        assert a["fitness_score"] == other["fitness_score"]
    if Oracle.available("nanoflann"):
        d = Oracle("nanoflann").icp_align(src, tgt, threads=4)
        assert np.array_equal(a["T"], d["T"]) and a["iterations"] == d["iterations"]


def test_oracle_icp_single_step_equals_numpy_umeyama(oracle, clouds):
    tgt = clouds
    src = apply(rigid(0.01, 0.0, -0.02, [0.2, 0.1, -0.05]), tgt[::4])
    r = oracle.icp_align(src, tgt, max_iterations=1, brute=True, threads=4)
    assert r["iterations"] == 1 and r["state"] == 1 and r["converged"] == 1       # iteration limit counts as converged
    # independent step: exhaustive nearest neighbours in f32, Kabsch in f64
    d = src[:, None, :3] - tgt[None, :, :3]
    d2 = (d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1]) + d[:, :, 2] * d[:, :, 2]
    j = d2.argmin(axis=1)
    S, D = src[:, :3].astype(np.float64), tgt[j, :3].astype(np.float64)
    ms, md = S.mean(0), D.mean(0)
    sigma = (D - md).T @ (S - ms) / S.shape[0]
    U, sv, Vt = np.linalg.svd(sigma)
    dd = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        dd[2] = -1
    R = U @ np.diag(dd) @ Vt
    t = md - R @ ms
    assert np.abs(r["T"][:3, :3] - R).max() < 2e-6 and np.abs(r["T"][:3, 3] - t).max() < 2e-4
    assert r["last_mse"] == pytest.approx(float(d2.min(axis=1).astype(np.float64).mean()), rel=1e-9)


def test_oracle_icp_max_distance_and_no_correspondences(oracle, clouds):
    tgt = clouds
    src = apply(rigid(0, 0, 0, [300.0, 0, 0]), tgt[::10])
    r = oracle.icp_align(src, tgt, max_correspondence_distance=5.0, threads=4)
    assert r["state"] == 5 and r["converged"] == 0 and r["iterations"] == 0
    assert np.array_equal(r["T"], np.eye(4, dtype=np.float32))
