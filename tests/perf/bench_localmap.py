#!/usr/bin/env python
"""publishLocalMap on 50 keyframes of voxelised 32-beam sweeps: device and wall time per call for the
settings of utility.h / jeep.yaml, checked bit-exact against the oracle first.  One JSON line per setting.
`--one` runs a single setting a few times (the command profiled by ncu for profiles/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lio_slam_b200 import synth  # noqa: E402
from lio_slam_b200.liogpu import LioGpu  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    one = "--one" in sys.argv
    o = Oracle("nanoflann") if Oracle.available("nanoflann") else Oracle("port")
    world = synth.make_world(1234)
    clouds, poses = [], []
    for k in range(50):
        p = synth.path_pose(0.5 * k)
        ds, _ = o.voxel_grid(synth.to_packed(synth.make_scan(world, p, 32, seed=700 + k, cols=900)), 0.4)
        clouds.append(ds)
        poses.append(p.astype(np.float32))
    poses = np.array(poses)
    g = LioGpu()
    ids = list(range(50))
    for k, c in enumerate(clouds):
        g.keyframe_put(k, c)
    settings = [("outlier filter meanK 10 + VoxelGrid leaf 0.2", dict(leaf=0.2), dict(local_mapping_surf_leaf_size=0.2))]
    if not one:
        settings += [("utility.h defaults (leaf 0.01: overflow guard)", dict(), dict()),
                     ("jeep.yaml: no outlier filter, leaf 0.2", dict(leaf=0.2, use_removing_outliers=False),
                      dict(local_mapping_surf_leaf_size=0.2, use_removing_outliers=0)),
                     ("meanK 30, 2 sigma, leaf 0.2", dict(leaf=0.2, mean_k=30, stddev_threshold=2.0),
                      dict(local_mapping_surf_leaf_size=0.2, mean_k=30, stddev_threshold=2.0))]
    for label, okw, gkw in settings:
        want, winfo, _ = o.publish_local_map(clouds, poses, poses[-1], threads=os.cpu_count(), **okw)
        dev, wall = [], []
        for _ in range(3 if one else 12):
            t0 = time.perf_counter()
            got, info, st = g.publish_local_map(ids, poses, poses[-1], **gkw)
            wall.append(1e3 * (time.perf_counter() - t0))
            dev.append(info["gpu_ms"])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        print(json.dumps(dict(setting=label, gpu_device_ms=float(np.median(dev[2:])), gpu_wall_ms=float(np.median(wall[2:])),
                              bit_equal=True, **{k: info[k] for k in ("n_concat", "n_cropped", "n_after_sor", "n_out",
                                                                      "sor_leftover", "sor_exhaustive", "leaf_overflow")})))
    g.close()


if __name__ == "__main__":
    main()
