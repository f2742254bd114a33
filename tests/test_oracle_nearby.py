"""CPU check of the oracle's extractNearby restatement (mapOptmization.cpp:1519-1565) against an
independent numpy version (scipy cKDTree for the radius and nearest searches, the oracle's own VoxelGrid for the
density filter, which test_oracle_core.py checks separately)."""
import numpy as np
from scipy.spatial import cKDTree


def np_extract_nearby(oracle, key3d, t, t_cur, radius, density):
    n = key3d.shape[0]
    last = key3d[-1]
    d = last[:3] - key3d[:, :3]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]                  # f32, L2_Simple order
    r2 = np.float32(np.float64(radius) * np.float64(radius))
    hits = np.nonzero(d2 < r2)[0]
    hits = hits[np.argsort(d2[hits], kind="stable")]
    ds, _ = oracle.voxel_grid(key3d[hits], density)
    tree = cKDTree(key3d[:, :3].astype(np.float64))
    entries = []
    for p in ds:
        # nearest key pose in f32 arithmetic, ties to the lower index: candidates from the f64 tree, decided in f32
        _, cand = tree.query(p[:3].astype(np.float64), k=min(8, n))
        cand = np.atleast_1d(cand)
        dd = p[:3] - key3d[cand, :3]
        c2 = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
        best = cand[np.lexsort((cand, c2))[0]]
        entries.append((p[:3], int(key3d[best, 3])))
    for i in range(n - 1, -1, -1):
        if t_cur - t[i] < 10.0:
            entries.append((key3d[i, :3], int(key3d[i, 3])))
        else:
            break
    ids = []
    for xyz, kid in entries:
        dd = xyz - last[:3]
        dist = np.float32(np.sqrt(np.float64(np.float32((dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2]))))
        if dist > np.float32(radius):
            continue
        ids.append(kid)
    return np.array(ids, np.int32)


def test_oracle_extract_nearby_vs_numpy(oracle):
    rng = np.random.default_rng(21)
    n = 1500
    ang = np.cumsum(rng.normal(0, 0.06, n))
    xyz = np.cumsum(np.c_[np.cos(ang), np.sin(ang), rng.normal(0, 0.01, n)] * 0.9, axis=0)
    key3d = np.c_[xyz, np.arange(n)].astype(np.float32)
    t = 0.45 * np.arange(n)
    for radius, density, dt in ((50.0, 2.0, 0.1), (20.0, 1.0, 0.1), (50.0, 2.0, 30.0), (8.0, 2.0, 0.1)):
        want = np_extract_nearby(oracle, key3d, t, t[-1] + dt, radius, density)
        got = oracle.extract_nearby(key3d, t, t[-1] + dt, radius, density)
        assert np.array_equal(got, want), (radius, density, dt)
    one = np.array([[1.0, 2.0, 3.0, 0.0]], np.float32)
    assert np.array_equal(oracle.extract_nearby(one, np.array([4.0]), 4.5), np.array([0, 0], np.int32))
