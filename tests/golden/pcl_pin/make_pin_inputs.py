#!/usr/bin/env python
"""Writes pin_inputs.lpin: the seeded inputs pcl_pin.cpp feeds to the real PCL / Eigen / OpenCV calls and
tests/test_pcl_pin.py feeds to the oracle.  Run from the repository root:  python tests/golden/pcl_pin/make_pin_inputs.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import lpin  # noqa: E402
from lio_slam_b200 import synth  # noqa: E402


def main():
    world = synth.make_world(1234)
    gt = synth.path_pose(0.0)
    a = synth.to_packed(synth.make_scan(world, gt, 16, seed=11, cols=450))            # ~7k points, sensor frame
    b = synth.make_local_map(world, 16, 12000, 0.5, seed=5, s0=-0.5, cols=900, max_poses=32)   # map frame
    guess = synth.perturbed_guess(gt, 21)
    rng = np.random.default_rng(7)
    # loop-closure pair: a keyframe-sized source, perturbed, against a submap
    src = synth.transform_packed(a[::2], synth.perturbed_guess(gt, 5, trans=(0.4, 0.3, 0.05)))
    # 6x6 normal equations as LMOptimization builds them: A (n x 6), b (n)
    A6 = rng.normal(0, 1, (400, 6)).astype(np.float32) * np.array([3, 3, 3, 1, 1, 1], np.float32)
    b6 = rng.normal(0, 0.05, 400).astype(np.float32)
    lpin.write(os.path.join(HERE, "pin_inputs.lpin"), {
        "cloud_a": a.astype(np.float32), "cloud_b": b.astype(np.float32), "pose_guess": guess.astype(np.float32),
        "icp_source": src.astype(np.float32), "lm_A": A6, "lm_b": b6,
        "pose_now": np.array([0.01, -0.02, 0.7, 3.5, -1.25, 0.4], np.float32),
    })
    print("wrote", os.path.join(HERE, "pin_inputs.lpin"), a.shape, b.shape, src.shape)


if __name__ == "__main__":
    main()
