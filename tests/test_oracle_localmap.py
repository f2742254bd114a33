"""CPU checks of the oracle's publishLocalMap restatement (mapOptmization.cpp:2442-2541):
its three (meanK+1)-NN providers agree, and the whole chain matches an independent numpy restatement written
from the PCL / Eigen semantics in the oracle header.  (Still PARITY UNPINNED: no PCL in this image.)"""
import numpy as np
import pytest

from lio_slam_b200 import synth


def np_transform(c, pose):
    """transformPointCloud (MO:849-868) in f32, products summed left to right."""
    roll, pitch, yaw = [np.float32(v) for v in pose[:3]]
    f = np.float32
    A, B = f(np.cos(np.float64(yaw))), f(np.sin(np.float64(yaw)))
    C_, D = f(np.cos(np.float64(pitch))), f(np.sin(np.float64(pitch)))
    E, F = f(np.cos(np.float64(roll))), f(np.sin(np.float64(roll)))
    DE, DF = f(D * E), f(D * F)
    T = np.array([[A * C_, A * DF - B * E, B * F + A * DE, pose[3]],
                  [B * C_, A * E + B * DF, B * DE - A * F, pose[4]],
                  [-D, C_ * F, C_ * E, pose[5]]], np.float32)
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    out = c.copy()
    for r in range(3):
        out[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
    return out


def np_publish(clouds, poses, now, left, right, front, back, sor, mean_k, std):
    f = np.float32
    cat = np.concatenate([np_transform(c, p) for c, p in zip(clouds, poses)]).astype(np.float32)
    nyaw = f(-now[2])
    c, s = f(np.cos(np.float64(nyaw))), f(np.sin(np.float64(nyaw)))
    tX = f(f(now[3] * c) - f(now[4] * s))
    tY = f(f(now[4] * c) + f(now[3] * s))
    zz = f(f(f(1.0) - c) + c)
    x, y, z = cat[:, 0], cat[:, 1], cat[:, 2]
    zero = f(0.0)
    q = cat.copy()
    q[:, 0] = x * c + (y * f(-s) + (z * zero + f(-tX)))
    q[:, 1] = x * s + (y * c + (z * zero + f(-tY)))
    q[:, 2] = x * zero + (y * zero + (z * zz + f(-now[5])))
    keep = np.isfinite(q[:, :3]).all(axis=1) & (q[:, 0] >= f(-left)) & (q[:, 0] <= f(right)) & \
        (q[:, 1] >= f(-back)) & (q[:, 1] <= f(front))
    q = q[keep]
    md = None
    if sor and q.shape[0] > mean_k:
        p = q[:, :3]
        d = p[:, None, :] - p[None, :, :]
        d2 = (d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1]) + d[:, :, 2] * d[:, :, 2]   # f32, L2_Simple order
        d2.sort(axis=1)
        root = np.sqrt(d2[:, 1:mean_k + 1])                    # f32 sqrt, neighbour 0 = the point itself
        md = (root.astype(np.float64).cumsum(axis=1)[:, -1] / mean_k).astype(np.float32)   # sequential f64 sum
        ssum = 0.0
        sq = 0.0
        for v in md:                                           # sequential, like PCL
            ssum += float(v)
            sq += float(np.float32(v * v))
        n = float(md.shape[0])
        mean = ssum / n
        var = (sq - ssum * ssum / n) / (n - 1.0)
        thr = mean + float(std) * np.sqrt(var)
        q = q[~(md.astype(np.float64) > thr)]
    return q, md   # the VoxelGrid step is checked on its own in test_oracle_core.py


@pytest.fixture(scope="module")
def case(world, oracle):
    clouds, poses = [], []
    for k in range(3):
        p = synth.path_pose(1.0 * k)
        ds, _ = oracle.voxel_grid(synth.to_packed(synth.make_scan(world, p, 16, seed=900 + k, cols=200)), 0.4)
        clouds.append(ds)
        poses.append(p.astype(np.float32))
    return clouds, np.array(poses, np.float32)


@pytest.mark.parametrize("mean_k,std,leaf,sor,ds", [(10, 1.0, 0.3, True, True), (4, 0.5, 0.5, True, False),
                                                     (10, 1.0, 0.3, False, True), (10, 1.0, 0.01, True, True)])
def test_oracle_publish_local_map_vs_numpy(oracle, case, mean_k, std, leaf, sor, ds):
    clouds, poses = case
    now = np.array([0.01, -0.02, 0.7, 1.5, -0.8, 0.1], np.float32)
    want, wmd = np_publish(clouds, poses, now, 40.0, 40.0, 70.0, 20.0, sor, mean_k, std)
    if ds and leaf > 0.05:
        want, ov = oracle.voxel_grid(want, leaf)
        assert not ov
    got, info, md = oracle.publish_local_map(clouds, poses, now, use_removing_outliers=sor, mean_k=mean_k,
                                             stddev_threshold=std, use_down_sampling=ds, leaf=leaf, brute=True, threads=4)
    if sor:
        assert np.array_equal(md.view(np.uint32), wmd.view(np.uint32))
    assert info["leaf_overflow"] == (1 if (ds and leaf <= 0.05) else 0)   # leaf 0.01: guard, cloud unchanged (q4)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_oracle_sor_knn_providers_agree(oracle, case):
    from oracle.oracle import Oracle
    clouds, poses = case
    now = poses[1]
    a, ia, mda = oracle.publish_local_map(clouds, poses, now, leaf=0.2, brute=True, threads=4)
    b, ib, mdb = oracle.publish_local_map(clouds, poses, now, leaf=0.2, brute=False, threads=4)
    assert ia == ib and np.array_equal(mda, mdb) and np.array_equal(a, b)
    if Oracle.available("nanoflann"):
        nf = Oracle("nanoflann")
        c, ic, mdc = nf.publish_local_map(clouds, poses, now, leaf=0.2, threads=4)
        assert ic == ia and np.array_equal(mda, mdc) and np.array_equal(a, c)
    # thread count never changes anything
    d, id_, mdd = oracle.publish_local_map(clouds, poses, now, leaf=0.2, threads=1)
    assert id_ == ia and np.array_equal(a, d)


def test_oracle_yaw_frame_matrix(oracle):
    for yaw in (0.0, 0.4, -1.3, 2.9, -3.1):
        now = np.array([0.3, -0.2, yaw, 5.0, -7.0, 1.25], np.float32)
        m = oracle.yaw_frame_T(now).reshape(3, 4)
        c, s = np.float32(np.cos(-np.float64(np.float32(yaw)))), np.float32(np.sin(-np.float64(np.float32(yaw))))
        assert m[0, 0] == c and m[1, 1] == c and m[0, 1] == -s and m[1, 0] == s
        assert m[2, 2] == np.float32(np.float32(1.0) - c) + c
        assert m[0, 2] == 0 and m[1, 2] == 0 and m[2, 0] == 0 and m[2, 1] == 0
        # the origin of the vehicle maps to (0, 0, 0) up to rounding: translation = -R(-yaw) * t
        v = m[:, :3].astype(np.float64) @ now[3:].astype(np.float64) + m[:, 3]
        assert np.abs(v).max() < 1e-5
