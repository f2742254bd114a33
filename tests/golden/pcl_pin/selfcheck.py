#!/usr/bin/env python
"""Dry run of tests/test_pcl_pin.py: writes an outputs file computed BY THE ORACLE ITSELF (not by the real libraries)
to a temporary path and runs the pin tests against it.  It proves nothing about parity — it only checks that the record
names, shapes and comparison code of the kit and of the tests fit together, so that the first real fixture does not
fail on plumbing.  Run from the repository root:  python tests/golden/pcl_pin/selfcheck.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import lpin  # noqa: E402
from oracle.oracle import Oracle, build  # noqa: E402


def main():
    build()
    o = Oracle("port")
    inp = lpin.read(os.path.join(HERE, "pin_inputs.lpin"))
    a04, _ = o.voxel_grid(inp["cloud_a"], 0.4)
    b05, _ = o.voxel_grid(inp["cloud_b"], 0.5)
    a20, _ = o.voxel_grid(inp["cloud_a"], 2.0)
    T = np.vstack([o.pose_to_T(inp["pose_guess"]).reshape(3, 4), [0, 0, 0, 1]]).astype(np.float32)
    res = o.surf_optimization(b05, a04, T12=T[:3].reshape(12), threads=4)
    qr = np.array([o.qr53_solve(b05[res["nn_idx"][i], :3], -np.ones(5, np.float32)) for i in range(a04.shape[0])], np.float32)
    lm, info, _ = o.publish_local_map([inp["cloud_b"]], np.zeros((1, 6), np.float32), inp["pose_now"], use_removing_outliers=True,
                                      mean_k=10, stddev_threshold=1.0, use_down_sampling=False, threads=4)
    icp = o.icp_align(inp["icp_source"], b05, threads=4)
    A, b = inp["lm_A"], inp["lm_b"]
    AtA = (A.astype(np.float64).T @ A.astype(np.float64)).astype(np.float32)
    Atb = (A.astype(np.float64).T @ b.astype(np.float64)).astype(np.float32)
    _, x = o.cv_solve6_qr(AtA, Atb)
    E, V = o.cv_eigen6(AtA)
    _, Vi = o.cv_inv6(V)
    n = inp["cloud_a"].shape[0]
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "self.lpin")
        lpin.write(out, {"vox_a_04": a04, "vox_b_05": b05, "vox_a_20": a20, "vox_guard_n": np.array([n, n], np.int32),
                         "T_pose": T, "knn_idx": res["nn_idx"].astype(np.int32), "knn_d2": res["nn_d2"], "qr_x": qr,
                         "sor_out": lm, "icp_T": icp["T"].astype(np.float32),
                         "icp_meta": np.array([float(icp["converged"]), icp["fitness_score"]], np.float64),
                         "cv_AtA": AtA, "cv_x": x.reshape(6, 1), "cv_E": E.reshape(1, 6), "cv_V": V, "cv_Vinv": Vi})
        env = dict(os.environ, PCL_PIN_OUT=out)
        return subprocess.call([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_pcl_pin.py"), "-q"], env=env, cwd=ROOT)


if __name__ == "__main__":
    sys.exit(main())
