"""GPU parity of liogpu_icp_align against the oracle's restatement of the loop-closure registration
(pcl::IterativeClosestPoint as configured at mapOptmization.cpp:1111-1123).
Floating-point outputs: final transformation within 1e-4 m / 1e-5 rad of the oracle (north_star's pose tolerance;
observed ~1e-7: the f64 reductions differ only in summation order), fitness score 1e-6 relative; integer outputs
(iteration count, convergence state, correspondence count) equal."""
import numpy as np
import pytest

from lio_slam_b200 import synth

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    """small-angle distance between two rotation matrices: |Ra - Rb|_F / sqrt(2) (arccos of the trace is useless here:
    f32 matrices are orthonormal only to 1e-7, which arccos turns into 4e-4)"""
    return float(np.linalg.norm(Ra.astype(np.float64) - Rb.astype(np.float64)) / np.sqrt(2.0))


def check(gpu, oracle, src, tgt, radius=10.0, brute=False, same_iterations=True, **over):
    okw = dict(max_correspondence_distance=over.get("max_correspondence_distance", 2 * radius),
               max_iterations=over.get("max_iterations", 100),
               transformation_epsilon=over.get("transformation_epsilon", 1e-6),
               euclidean_fitness_epsilon=over.get("euclidean_fitness_epsilon", 1e-6))
    want = oracle.icp_align(src, tgt, brute=brute, threads=8, **okw)
    T, info = gpu.icp_align(src, tgt, radius, **over)
    assert info["converged"] == want["converged"] and info["convergence_state"] == want["state"]
    if same_iterations:
        assert info["iterations"] == want["iterations"]
    assert info["n_correspondences"] == want["n_correspondences"]
    assert np.abs(T[:3, 3] - want["T"][:3, 3]).max() < 1e-4                      # metres
    assert rot_angle(T[:3, :3], want["T"][:3, :3]) < 1e-5                         # radians
    assert np.array_equal(T[3], np.array([0, 0, 0, 1], np.float32))
    if want["n_correspondences"] >= 3:
        assert info["fitness_score"] == pytest.approx(want["fitness_score"], rel=1e-6, abs=1e-12)
        assert info["last_mse"] == pytest.approx(want["last_mse"], rel=1e-6, abs=1e-12)
    return T, info, want


@pytest.fixture(scope="module")
def loop_case(world, oracle):
    """cureKeyframeCloud / prevKeyframeCloud as loopFindNearKeyframes builds them (mapOptmization.cpp:1102-1103):
    one keyframe at a drifted pose against a submap of 11 neighbouring keyframes, all voxelised at 0.4 m"""
    clouds, poses = [], []
    for k in range(11):
        p = synth.path_pose(1.0 * k)
        ds, _ = oracle.voxel_grid(synth.to_packed(synth.make_scan(world, p, 16, seed=800 + k, cols=900)), 0.4)
        clouds.append(ds)
        poses.append(p.astype(np.float32))
    poses = np.array(poses, np.float32)
    p = synth.path_pose(5.3)
    cur, _ = oracle.voxel_grid(synth.to_packed(synth.make_scan(world, p, 16, seed=899, cols=900)), 0.4)
    wrong = p.astype(np.float32).copy()
    wrong[3] += 0.8
    wrong[4] -= 0.5
    wrong[2] += 0.03
    return dict(clouds=clouds, poses=poses, cur=cur, wrong=wrong, true=p.astype(np.float32))


def test_icp_loop_closure_parity(gpu, oracle, loop_case):
    ids = []
    gpu.keyframe_clear()
    for k, c in enumerate(loop_case["clouds"]):
        gpu.keyframe_put(k, c)
        ids.append(k)
    gpu.keyframe_put(100, loop_case["cur"])
    # loopFindNearKeyframes: transform + concatenate + downSizeFilterICP
    tgt, _ = gpu.merge_keyframes(ids, loop_case["poses"], 0.4)
    src, _ = gpu.merge_keyframes([100], loop_case["wrong"].reshape(1, 6), 0.4)
    want_tgt, _ = oracle.build_local_map(loop_case["clouds"], loop_case["poses"], 0.4, threads=4)
    assert np.array_equal(tgt.view(np.uint32), want_tgt.view(np.uint32))
    assert src.shape[0] >= 300 and tgt.shape[0] >= 1000            # the guard of mapOptmization.cpp:1104
    T, info, want = check(gpu, oracle, src, tgt, radius=10.0)
    assert info["converged"] == 1 and info["fitness_score"] < 0.3   # historyKeyframeFitnessScore
    # the correction undoes the injected drift: correctionLidarFrame * tWrong ~ true pose (mapOptmization.cpp:1138-1142)
    Tw = np.eye(4)
    Tw[:3, :3] = synth.rpy_to_R(*[float(v) for v in loop_case["wrong"][:3]])
    Tw[:3, 3] = loop_case["wrong"][3:]
    corrected = T.astype(np.float64) @ Tw
    assert np.abs(corrected[:3, 3] - loop_case["true"][3:]).max() < 0.1


def test_icp_matches_exhaustive_search(gpu, oracle, loop_case):
    tgt, _ = oracle.build_local_map(loop_case["clouds"][3:8], loop_case["poses"][3:8], 0.6, threads=4)
    src = oracle.transform_cloud(loop_case["cur"][::3], loop_case["wrong"])
    check(gpu, oracle, src, tgt, radius=15.0, brute=True)


@pytest.mark.parametrize("maxd", [0.3, 1.0, 3.0])
def test_icp_limited_correspondence_distance(gpu, oracle, loop_case, maxd):
    tgt, _ = oracle.build_local_map(loop_case["clouds"], loop_case["poses"], 0.4, threads=4)
    src = oracle.transform_cloud(loop_case["cur"], loop_case["wrong"])
    T, info, want = check(gpu, oracle, src, tgt, max_correspondence_distance=maxd)
    assert 3 <= info["n_correspondences"] <= src.shape[0]


def test_icp_iteration_limit_and_identity(gpu, oracle, loop_case):
    tgt, _ = oracle.build_local_map(loop_case["clouds"], loop_case["poses"], 0.4, threads=4)
    src = oracle.transform_cloud(loop_case["cur"], loop_case["wrong"])
    T, info, _ = check(gpu, oracle, src, tgt, max_iterations=3)
    assert info["iterations"] == 3 and info["convergence_state"] == 1 and info["converged"] == 1
    # a source that is a subset of the target: every nearest neighbour is the point itself
    T, info, _ = check(gpu, oracle, tgt[::7].copy(), tgt)
    assert info["iterations"] == 1 and info["fitness_score"] == 0.0
    assert np.abs(T - np.eye(4, dtype=np.float32)).max() < 1e-6


def test_icp_no_correspondences_and_bad_arguments(gpu, oracle, loop_case):
    from lio_slam_b200.liogpu import LioGpuError
    tgt, _ = oracle.build_local_map(loop_case["clouds"][:3], loop_case["poses"][:3], 0.4, threads=4)
    far = loop_case["cur"].copy()
    far[:, 0] += 500.0
    T, info, want = check(gpu, oracle, far, tgt, max_correspondence_distance=5.0)
    assert info["convergence_state"] == 5 and info["converged"] == 0 and info["iterations"] == 0
    assert np.array_equal(T, np.eye(4, dtype=np.float32))
    with pytest.raises(LioGpuError):
        gpu.icp_align(np.zeros((0, 4), np.float32), tgt)
    with pytest.raises(LioGpuError):
        gpu.icp_align(far, tgt, max_iterations=0)


@pytest.mark.parametrize("cell", [0.3, 2.5, 8.0])
def test_icp_independent_of_cell_size(gpu, oracle, loop_case, cell):
    tgt, _ = oracle.build_local_map(loop_case["clouds"], loop_case["poses"], 0.4, threads=4)
    src = oracle.transform_cloud(loop_case["cur"], loop_case["wrong"])
    T0, i0 = gpu.icp_align(src, tgt)
    T1, i1 = gpu.icp_align(src, tgt, cell_size=cell)
    assert np.array_equal(T0.view(np.uint32), T1.view(np.uint32))
    for k in ("iterations", "convergence_state", "n_correspondences", "fitness_score", "last_mse"):
        assert i0[k] == i1[k], k


def test_icp_sparse_target_and_outliers(gpu, oracle):
    """isolated source points far from a sparse target exercise the wide and the exhaustive search tiers"""
    rng = np.random.default_rng(3)
    tgt = np.c_[rng.uniform(-40, 40, (3000, 2)), rng.normal(0, 0.3, (3000, 1)), np.zeros((3000, 1))].astype(np.float32)
    src = tgt[rng.choice(3000, 1200, replace=False)].copy()
    src[:, 0] += 0.4
    src[:, 1] -= 0.2
    stray = np.c_[rng.uniform(-40, 40, (60, 2)), rng.uniform(10, 18, (60, 1)), np.zeros((60, 1))].astype(np.float32)
    src = np.vstack([src, stray])
    check(gpu, oracle, src, tgt, radius=10.0, brute=True)
    check(gpu, oracle, src, tgt, max_correspondence_distance=4.0, brute=True)


def test_icp_keeps_registration_index(gpu, oracle, small_case, loop_case):
    gpu.set_local_map(small_case["map4"])
    p0, _, i0 = gpu.scan2map(small_case["scan4"], small_case["guess"])
    tgt, _ = oracle.build_local_map(loop_case["clouds"][:4], loop_case["poses"][:4], 0.4, threads=4)
    gpu.icp_align(oracle.transform_cloud(loop_case["cur"], loop_case["wrong"]), tgt)
    p1, _, i1 = gpu.scan2map(small_case["scan4"], small_case["guess"])
    assert np.array_equal(p0.view(np.uint32), p1.view(np.uint32)) and i0["iterations"] == i1["iterations"]
