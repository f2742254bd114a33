"""CPU check of the oracle's Scan Context restatement (include/Scancontext.cpp:23-36, 151-225)
against an independent vectorised numpy version of the same source lines."""
import numpy as np

from lio_slam_b200 import synth


def np_scancontext(c, lidar_height=2.0, max_radius=80.0):
    c = c[np.isfinite(c[:, :3]).all(axis=1)]
    x, y = c[:, 0], c[:, 1]
    z = (c[:, 2].astype(np.float64) + lidar_height).astype(np.float32)
    rng = np.sqrt((x * x + y * y).astype(np.float64)).astype(np.float32)
    k = 180.0 / np.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        th = np.where((x >= 0) & (y >= 0), k * np.arctan((y / x).astype(np.float64)),
             np.where((x < 0) & (y >= 0), 180 - k * np.arctan((y / (-x)).astype(np.float64)),
             np.where((x < 0) & (y < 0), 180 + k * np.arctan((y / x).astype(np.float64)),
                      360 - k * np.arctan(((-y) / x).astype(np.float64))))).astype(np.float32)
    keep = ~(rng.astype(np.float64) > max_radius)
    z, rng, th = z[keep], rng[keep], th[keep]
    ring = np.clip(np.ceil(rng.astype(np.float64) / max_radius * 20), 1, 20).astype(int)
    sec = np.ceil(th.astype(np.float64) / 360.0 * 60)
    sec = np.clip(np.where(np.isnan(sec), 0, sec), 1, 60).astype(int)
    desc = np.full((20, 60), -1000.0)
    np.maximum.at(desc, (ring - 1, sec - 1), z.astype(np.float64))
    desc[desc == -1000.0] = 0.0
    return desc, desc.sum(axis=1) / 60, desc.sum(axis=0) / 20


def test_oracle_scancontext_vs_numpy(oracle, world):
    for beams, seed in ((16, 3), (64, 4)):
        scan = synth.to_packed(synth.make_scan(world, synth.path_pose(1.0 * seed), beams, seed=seed, cols=900))
        for kw in (dict(), dict(lidar_height=0.0, max_radius=35.0)):
            want = np_scancontext(scan, **kw)
            got = oracle.make_scancontext(scan, **kw)
            assert np.array_equal(got[0], want[0])
            assert np.allclose(got[1], want[1], rtol=0, atol=1e-12) and np.allclose(got[2], want[2], rtol=0, atol=1e-12)
    odd = np.array([[0, 0, 1, 0], [0, 3, 1, 0], [-3, 0, 2, 0], [80, 0, 5, 0], [80.001, 0, 9, 0], [np.nan, 1, 1, 0],
                    [1, 1, -1500, 0]], np.float32)
    got, want = oracle.make_scancontext(odd), np_scancontext(odd)
    assert np.array_equal(got[0], want[0]) and got[0][19, 59] == 0 and got[0].max() == 7.0
