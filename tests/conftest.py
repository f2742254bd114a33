"""pytest configuration: the `gpu` marker, repo-root imports and shared synthetic fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build()
    return Oracle("port")


@pytest.fixture(scope="session")
def world():
    from lio_slam_b200 import synth
    return synth.make_world(1234)


@pytest.fixture(scope="session")
def small_case(world):
    """16-beam / 900-column sweep, 12k-point local map (leaf 0.5), perturbed initial guess."""
    from lio_slam_b200 import synth
    pose_gt = synth.path_pose(0.0)
    scan = synth.make_scan(world, pose_gt, 16, seed=11, cols=900)
    map4 = synth.make_local_map(world, 16, 12000, 0.5, seed=5, s0=-0.5, cols=900, max_poses=32)
    guess = synth.perturbed_guess(pose_gt, 21)
    return dict(pose_gt=pose_gt, scan=scan, scan4=synth.to_packed(scan), map4=map4, guess=guess)


@pytest.fixture(scope="session")
def gpu():
    from lio_slam_b200.liogpu import LioGpu
    g = LioGpu()
    yield g
    g.close()
