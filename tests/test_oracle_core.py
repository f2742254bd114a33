"""CPU checks of the oracle itself (no GPU): VoxelGrid semantics, exact k-NN (brute force vs the two
KD-trees), plane fit against numpy, the committed end-to-end golden case, and deskew invariants."""
import os

import numpy as np
import pytest

from lio_slam_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "path_small.npz"))


def biteq(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_voxel_grid_semantics(oracle):
    rng = np.random.default_rng(0)
    cloud = np.column_stack([rng.uniform(-20, 20, 20000), rng.uniform(-20, 20, 20000), rng.uniform(-2, 2, 20000),
                             rng.uniform(0, 100, 20000)]).astype(np.float32)
    leaf = 0.5
    out, ov = oracle.voxel_grid(cloud, leaf)
    assert not ov
    inv = np.float32(1) / np.float32(leaf)
    ijk = np.floor(cloud[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(axis=0)
    div = ijk.max(axis=0) + 1
    key = ijk[:, 0] + div[0] * (ijk[:, 1] + div[1] * ijk[:, 2])
    uniq, first = np.unique(key, return_index=True)
    assert out.shape[0] == uniq.shape[0]                       # one output per occupied voxel, ascending index
    for v in (0, 17, uniq.shape[0] - 1):                       # sequential f32 sums in input order, true division
        members = np.flatnonzero(key == uniq[v])
        s = np.zeros(4, np.float32)
        for m in members:
            s = (s + cloud[m]).astype(np.float32)
        assert biteq(out[v], s / np.float32(members.size))
    out2, _ = oracle.voxel_grid(out, leaf)                     # idempotent on its own output
    assert biteq(out2, out)
    one, _ = oracle.voxel_grid(cloud[:1], leaf)
    assert biteq(one, cloud[:1])
    empty, ov = oracle.voxel_grid(np.zeros((0, 4), np.float32), leaf)
    assert empty.shape == (0, 4) and not ov
    same, ov = oracle.voxel_grid(cloud, 0.0005)                # overflow guard: input returned unchanged (q4)
    assert ov and biteq(same, cloud)


def test_knn_kdtree_equals_bruteforce(oracle):
    rng = np.random.default_rng(2)
    map4 = rng.uniform(-10, 10, (30000, 4)).astype(np.float32)
    map4[:, 2] *= 0.1
    q = rng.uniform(-10, 10, (3000, 4)).astype(np.float32)
    q[:, 2] *= 0.1
    bi, bd, bt = oracle.knn5(map4, q, threads=8)
    h = oracle.index_build(map4)
    ki, kd, kt = oracle.knn5(map4, q, handle=h, threads=8)
    oracle.index_free(h)
    assert np.array_equal(bi, ki) and biteq(bd, kd) and np.array_equal(bt, kt)
    assert (np.diff(bd, axis=1) >= 0).all()
    # distances are FLANN L2_Simple in f32: ((dx*dx)+dy*dy)+dz*dz
    j = 5
    p = map4[bi[j, 0]]
    d = (q[j, :3] - p[:3]).astype(np.float32)
    r = np.float32(0) + d[0] * d[0]; r = np.float32(r + d[1] * d[1]); r = np.float32(r + d[2] * d[2])
    assert np.float32(r) == bd[j, 0]


def test_knn_nanoflann_reference_tree_agrees(oracle):
    from oracle.oracle import Oracle
    if not Oracle.available("nanoflann"):
        pytest.skip("oracle/_ref (reference-vendored nanoflann) not built")
    nf = Oracle("nanoflann")
    rng = np.random.default_rng(3)
    map4 = rng.uniform(-10, 10, (20000, 4)).astype(np.float32)
    q = rng.uniform(-10, 10, (2000, 4)).astype(np.float32)
    bi, bd, bt = oracle.knn5(map4, q, threads=8)
    h = nf.index_build(map4)
    ni, nd, nt = nf.knn5(map4, q, handle=h, threads=8)
    nf.index_free(h)
    assert biteq(bd, nd)
    notie = bt == 0
    assert np.array_equal(bi[notie], ni[notie])


def test_plane_fit_against_numpy(oracle):
    rng = np.random.default_rng(4)
    for _ in range(200):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        d = rng.uniform(1, 30)
        basis = np.linalg.svd(n.reshape(1, 3))[2][1:]
        pts = (rng.uniform(-0.5, 0.5, (5, 2)) @ basis - d * n + rng.normal(0, 1e-3, (5, 3))).astype(np.float32)
        x = oracle.qr53_solve(pts, -np.ones(5, np.float32))
        ref = np.linalg.lstsq(pts.astype(np.float64), -np.ones(5), rcond=None)[0]
        assert np.abs(x - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


def test_pose_to_T_matches_float64(oracle):
    rng = np.random.default_rng(5)
    for _ in range(50):
        pose = np.concatenate([rng.uniform(-0.3, 0.3, 2), rng.uniform(-3.1, 3.1, 1), rng.uniform(-50, 50, 3)]).astype(np.float32)
        T = oracle.pose_to_T(pose).reshape(3, 4)
        R = synth.rpy_to_R(*[float(v) for v in pose[:3]])
        assert np.abs(T[:, :3] - R).max() < 3e-7 and np.array_equal(T[:, 3], pose[3:])


def test_golden_path_small(oracle):
    ds, ov = oracle.voxel_grid(GOLD["scan4"], 0.4)
    assert biteq(ds, GOLD["ds"])
    surf = oracle.surf_optimization(GOLD["map4"], ds, pose6=GOLD["guess"], threads=4)
    assert np.array_equal(surf["nn_idx"], GOLD["nn_idx"]) and biteq(surf["nn_d2"], GOLD["nn_d2"])
    assert np.array_equal(surf["flag"], GOLD["flag"]) and biteq(surf["coeff"], GOLD["coeff"])
    h = oracle.index_build(GOLD["map4"])
    pose, P, info = oracle.scan2map(GOLD["map4"], ds, GOLD["guess"], threads=4, handle=h)
    oracle.index_free(h)
    assert info["iterations"] == int(GOLD["iterations"]) and np.array_equal(info["nsel_hist"], GOLD["nsel_hist"])
    assert biteq(pose, GOLD["pose"]) and biteq(info["pose_hist"], GOLD["pose_hist"])


def test_scan2map_threads_and_backends_agree(oracle):
    ds = GOLD["ds"]
    a = oracle.scan2map(GOLD["map4"], ds, GOLD["guess"], threads=1, brute=True)
    b = oracle.scan2map(GOLD["map4"], ds, GOLD["guess"], threads=8)
    assert biteq(a[0], b[0]) and a[2]["iterations"] == b[2]["iterations"]


def test_deskew_invariants(oracle, world):
    from oracle.oracle import DeskewParams
    scan = synth.make_scan(world, synth.path_pose(1.0), 32, seed=77, cols=300)
    t0 = 100.5
    imu_t, rx, ry, rz = synth.make_imu_table(t0, seed=8)
    dp = DeskewParams(32, 1, 1, 2.0, 10.0, 2.0, 2.0, 100.0, 90.0)
    out = oracle.deskew(scan, dp, t0, imu_t, rx, ry, rz, True)
    raw = synth.to_packed(scan)
    keep = ~((raw[:, 1] < 2.0) & (-10.0 < raw[:, 1]) & (raw[:, 0] < 2.0) & (-2.0 < raw[:, 0])) & (raw[:, 3] <= 90.0) \
        & (np.sqrt((raw[:, :3] ** 2).sum(1)) <= 100.0)
    assert out.shape[0] == int(keep.sum())
    # pure rotation about the sensor origin: ranges are preserved, intensity untouched, order kept
    kept = raw[keep]
    assert np.abs(np.linalg.norm(out[:, :3], axis=1) - np.linalg.norm(kept[:, :3], axis=1)).max() < 2e-4
    assert np.array_equal(out[:, 3], kept[:, 3])
    # the first surviving point defines the start frame: it maps to itself (q6)
    assert np.abs(out[0, :3] - kept[0, :3]).max() < 1e-5
    # decimation uses the RAW index (q5)
    dp2 = DeskewParams(32, 2, 3, 2.0, 10.0, 2.0, 2.0, 100.0, 90.0)
    out2 = oracle.deskew(scan, dp2, t0, imu_t, rx, ry, rz, False)
    idx = np.arange(raw.shape[0])
    keep2 = keep & (scan["ring"] % 2 == 0) & (idx % 3 == 0)
    assert biteq(out2, raw[keep2])


def test_knn_against_scipy_ckdtree(oracle):
    # an independent exact k-NN (scipy cKDTree, f64): same neighbour sets wherever the 5th/6th distances are
    # not within f32 rounding of each other, same order, f32 distances equal to the f64 ones within 1 ulp-ish
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(21)
    map4 = rng.uniform(-15, 15, (40000, 4)).astype(np.float32)
    map4[:, 2] *= 0.05
    q = rng.uniform(-15, 15, (5000, 4)).astype(np.float32)
    q[:, 2] *= 0.05
    idx, d2, tie = oracle.knn5(map4, q, threads=8)
    dd, ii = cKDTree(map4[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=6)
    clear = (dd[:, 5] - dd[:, 4]) > 1e-5 * np.maximum(dd[:, 5], 1e-3)
    inner = np.all(np.diff(dd[:, :5], axis=1) > 1e-5 * np.maximum(dd[:, 1:5], 1e-3), axis=1)
    ok = clear & inner
    assert ok.mean() > 0.95
    assert np.array_equal(idx[ok], ii[ok, :5])
    assert np.abs(np.sqrt(d2[ok].astype(np.float64)) - dd[ok, :5]).max() < 1e-5


def test_golden_rows_f_small(oracle):
    """the committed small cases of the widened rows (publishLocalMap, merging, ICP, Scan Context, key-pose selection):
    the oracle still produces exactly what tests/golden/make_golden.py stored"""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "rows_f_small.npz"))
    offs = G["offsets"]
    clouds = [G["clouds"][offs[k]:offs[k + 1]] for k in range(len(offs) - 1)]
    lm, info, md = oracle.publish_local_map(clouds, G["poses"], G["pose_now"], leaf=0.3, threads=2)
    assert np.array_equal(lm.view(np.uint32), G["local_map"].view(np.uint32))
    assert [info["n_concat"], info["n_cropped"], info["n_after_sor"], info["n_out"]] == G["local_map_counts"].tolist()
    assert np.array_equal(md.view(np.uint32), G["mean_distances"].view(np.uint32))
    merged, _ = oracle.build_local_map(clouds, G["poses"], 0.4, threads=2)
    assert np.array_equal(merged.view(np.uint32), G["merged"].view(np.uint32))
    icp = oracle.icp_align(G["icp_source"], merged, threads=2)
    assert np.array_equal(icp["T"].view(np.uint32), G["icp_T"].view(np.uint32))
    assert [icp["iterations"], icp["converged"], icp["state"], icp["n_correspondences"]] == G["icp_ints"].tolist()
    assert icp["fitness_score"] == float(G["icp_fitness"])
    sc, rk, sk = oracle.make_scancontext(clouds[0])
    assert np.array_equal(sc, G["sc_desc"]) and np.array_equal(rk, G["sc_ringkey"]) and np.array_equal(sk, G["sc_sectorkey"])
    ids = oracle.extract_nearby(G["key3d"], G["key_time"], G["key_time"][-1] + 0.1, 30.0, 2.0)
    assert np.array_equal(ids, G["nearby_ids"])
