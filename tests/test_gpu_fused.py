"""The one-launch LM loop (s2m_fused.cuh, params.s2m_path = 2) against the CPU oracle, iteration by iteration, and
against the two-kernel path (s2m_path = 1).  What is specific to it and therefore checked here:
  * the exact no-search certificate: every iteration's per-point neighbours / distances / coefficients / flags / tie
    bits must equal a surfOptimization pass of the oracle at the pose that iteration started from, whether the point
    was searched or certified, and the certificate must actually be taken on late iterations;
  * switching the certificate off, or taking the two-kernel path, must not change a single bit of the pose history;
  * dense maps (phase-1 gate + leftovers) and sparse maps (everything through the leftover search on iteration 0)."""
import numpy as np
import pytest

from lio_slam_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def dense_case(world):
    """32-beam sweep vs a 60k-point map at leaf 0.2: the phase-1 gate (0.4 m) is active, a few percent leftovers."""
    pose_gt = synth.path_pose(1.0)
    scan4 = synth.to_packed(synth.make_scan(world, pose_gt, 32, seed=77, cols=900))
    map4 = synth.make_local_map(world, 32, 60000, 0.2, seed=9, s0=0.5, cols=900, max_poses=16)
    guess = synth.perturbed_guess(pose_gt, 33)
    return dict(scan4=scan4, map4=map4, guess=guess, leaf=0.2)


def make_ctx(leaf, **over):
    from lio_slam_b200.liogpu import LioGpu
    return LioGpu(surrounding_keyframe_map_leaf_size=leaf, **over)


def run_all(case, scan, **over):
    g = make_ctx(case.get("leaf", 0.5), **over)
    try:
        g.set_local_map(case["map4"])
        return g.scan2map(scan, case["guess"])
    finally:
        g.close()


@pytest.mark.parametrize("which", ["small", "dense"])
def test_fused_equals_two_kernel_path_and_certificate_is_neutral(oracle, small_case, dense_case, which):
    case = dict(small_case, leaf=0.5) if which == "small" else dense_case
    scan = oracle.voxel_grid(case["scan4"], 0.4)[0] if which == "small" else case["scan4"]
    pose_f, P_f, info_f = run_all(case, scan, s2m_path=2)
    pose_n, P_n, info_n = run_all(case, scan, s2m_path=2, s2m_no_certificate=1)
    pose_l, P_l, info_l = run_all(case, scan, s2m_path=1)
    assert info_f["kernel_launches"] == 1 and info_l["kernel_launches"] > 2
    assert info_f["iterations"] == info_n["iterations"] == info_l["iterations"] >= 3
    # certificate on/off: the same neighbours in the same chunks -> the same sums in the same order
    assert np.array_equal(bits(info_f["pose_hist"]), bits(info_n["pose_hist"]))
    assert np.array_equal(info_f["JtJ"], info_n["JtJ"]) and np.array_equal(info_f["Jtr"], info_n["Jtr"])
    assert info_f["tie_queries"] == info_n["tie_queries"] == info_l["tie_queries"]
    assert np.array_equal(info_f["nsel_hist"], info_n["nsel_hist"]) and np.array_equal(info_f["nsel_hist"], info_l["nsel_hist"])
    # two-kernel path: another (fixed) order of the FP64 additions; JtJ agrees to ~1e-15, poses to the last bits
    scale = np.abs(info_l["JtJ"]).max()
    assert np.abs(info_f["JtJ"] - info_l["JtJ"]).max() <= 1e-12 * scale
    assert np.abs(info_f["pose_hist"] - info_l["pose_hist"]).max() <= 1e-6
    assert np.abs(P_f - P_l).max() <= 1e-5 and info_f["is_degenerate"] == info_l["is_degenerate"]
    assert info_f["certified"] > 0 and info_n["certified"] == 0
    print(f"{which}: iterations={info_f['iterations']} n={info_f['n_query']} certified(last)={info_f['certified']} "
          f"seeded(last)={info_f['seeded']} leftovers(last)={info_f['leftovers']} "
          f"pose bit-equal to the two-kernel path: {np.array_equal(bits(pose_f), bits(pose_l))}")


@pytest.mark.parametrize("which", ["small", "dense"])
def test_every_iteration_matches_the_oracle_point_by_point(oracle, small_case, dense_case, which):
    case = dict(small_case, leaf=0.5) if which == "small" else dense_case
    scan = oracle.voxel_grid(case["scan4"], 0.4)[0] if which == "small" else case["scan4"]
    g = make_ctx(case["leaf"], s2m_path=2)
    try:
        g.set_local_map(case["map4"])
        pose, P, info = g.scan2map(scan, case["guess"])
        iters = info["iterations"]
        cert_total = 0
        for k in range(1, iters + 1):
            pose_k, _, info_k, pts = g.scan2map_trace(scan, case["guess"], max_iter=k)
            assert info_k["iterations"] == k
            assert np.array_equal(bits(info_k["pose_hist"]), bits(info["pose_hist"][:k]))
            start = case["guess"] if k == 1 else info["pose_hist"][k - 2]
            ref = oracle.surf_optimization(case["map4"], scan, pose6=start, threads=8)
            gate = ref["nn_d2"][:, 4] < 1.0
            assert np.array_equal(pts["nn_idx"][gate], ref["nn_idx"][gate]), f"iteration {k}: neighbour sets"
            assert np.array_equal(bits(pts["nn_d2"][gate]), bits(ref["nn_d2"][gate])), f"iteration {k}: distances"
            assert np.array_equal(pts["nn_idx"][~gate], np.full((int((~gate).sum()), 5), -1)), f"iteration {k}: not-found rows"
            assert np.array_equal(pts["flag"], ref["flag"]), f"iteration {k}: flags"
            assert np.array_equal(bits(pts["coeff"]), bits(ref["coeff"])), f"iteration {k}: coefficients"
            assert np.array_equal(pts["tie"][gate], ref["tie"][gate]), f"iteration {k}: tie bits"
            cert_total += info_k["certified"]
            print(f"{which} iteration {k}: found={int(gate.sum())} accepted={int(ref['flag'].sum())} "
                  f"seeded={info_k['seeded']} certified={info_k['certified']} leftovers={info_k['leftovers']}")
        assert cert_total > 0, "the certificate was never taken"
    finally:
        g.close()


def test_fused_sparse_map_far_from_origin(oracle, small_case):
    # large coordinates stress the rounding margins of the certificate (they are relative, not absolute)
    off = np.array([4000.0, -7000.0, 300.0, 0.0], np.float32)
    map4 = (small_case["map4"] + off).astype(np.float32)
    ds, _ = oracle.voxel_grid(small_case["scan4"], 0.6)
    guess = small_case["guess"].copy(); guess[3:] += off[:3]
    ref_pose, ref_P, ref_info = oracle.scan2map(map4, ds, guess, threads=8)
    g = make_ctx(0.5, s2m_path=2)
    try:
        g.set_local_map(map4)
        pose, P, info = g.scan2map(ds, guess)
        assert info["iterations"] == ref_info["iterations"]
        assert np.array_equal(info["nsel_hist"], ref_info["nsel_hist"])
        assert info["tie_queries"] == ref_info["tie_queries"]
        assert np.abs(pose[:3] - ref_pose[:3]).max() <= 1e-5 and np.abs(pose[3:] - ref_pose[3:]).max() <= 1e-4 * 40
    finally:
        g.close()


def test_fused_is_reproducible_run_to_run(dense_case):
    # dynamic chunk queue, fixed summation order: two runs must agree bit for bit
    a = run_all(dense_case, dense_case["scan4"], s2m_path=2)
    b = run_all(dense_case, dense_case["scan4"], s2m_path=2)
    assert np.array_equal(bits(a[0]), bits(b[0]))
    assert np.array_equal(a[2]["JtJ"], b[2]["JtJ"]) and np.array_equal(bits(a[2]["pose_hist"]), bits(b[2]["pose_hist"]))


def test_fused_sequence_replay_matches_two_kernel_path(world):
    # the whole per-scan path (host mirror) with either LM loop: identical iteration counts, poses equal to the last bits
    import torch
    from lio_slam_b200 import replay, synth_torch
    seq = synth_torch.make_sequence(world, 64, 24, seed=9, device=torch.device("cuda", 0), step=0.35)
    a = replay.replay_sequence(replay.kitti_params(s2m_path=1), seq)
    b = replay.replay_sequence(replay.kitti_params(s2m_path=2), seq)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.abs(a[0] - b[0]).max() <= 2e-6
    print("bit-equal poses:", int((a[0].view(np.uint32) == b[0].view(np.uint32)).all(axis=1).sum()), "of 24")
