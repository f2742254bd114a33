#!/usr/bin/env python
"""Loop-closure ICP: one keyframe against a submap of 2*25+1 keyframes (historyKeyframeSearchNum 25),
both voxelised at 0.4 m, drifted by 0.9 m / 0.03 rad.  Device and wall time of liogpu_icp_align next to the CPU
oracle (KD-tree nearest neighbour, all host threads and one thread); outputs compared first."""
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
    o = Oracle("nanoflann") if Oracle.available("nanoflann") else Oracle("port")
    world = synth.make_world(1234)
    clouds, poses = [], []
    for k in range(51):
        p = synth.path_pose(0.5 * k)
        ds, _ = o.voxel_grid(synth.to_packed(synth.make_scan(world, p, 32, seed=600 + k, cols=900)), 0.4)
        clouds.append(ds)
        poses.append(p.astype(np.float32))
    tgt, _ = o.build_local_map(clouds, np.array(poses), 0.4, threads=os.cpu_count())
    p = synth.path_pose(12.7)
    cur, _ = o.voxel_grid(synth.to_packed(synth.make_scan(world, p, 32, seed=699, cols=900)), 0.4)
    wrong = p.astype(np.float32).copy()
    wrong[3] += 0.7; wrong[4] -= 0.6; wrong[2] += 0.03
    src = o.transform_cloud(cur, wrong)
    g = LioGpu()
    t0 = time.perf_counter(); want = o.icp_align(src, tgt, threads=os.cpu_count()); cpu_all = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); o.icp_align(src, tgt, threads=1); cpu_1 = 1e3 * (time.perf_counter() - t0)
    dev, wall = [], []
    for _ in range(8):
        t0 = time.perf_counter()
        T, info = g.icp_align(src, tgt)
        wall.append(1e3 * (time.perf_counter() - t0)); dev.append(info["gpu_ms"])
    assert info["iterations"] == want["iterations"] and np.abs(T - want["T"]).max() < 1e-4
    cells = {}
    for cell in (0.5, 0.75, 1.0, 1.5):
        d = []
        for _ in range(5):
            Tc, ic = g.icp_align(src, tgt, cell_size=cell)
            d.append(ic["gpu_ms"])
        assert np.array_equal(Tc, T)
        cells[str(cell)] = float(np.median(d[1:]))
    print(json.dumps(dict(stage="loop-closure ICP (f3)", gpu_device_ms_by_cell_size=cells, n_source=int(src.shape[0]), n_target=int(tgt.shape[0]),
                          iterations=info["iterations"], converged=info["converged"], fitness=info["fitness_score"],
                          gpu_device_ms=float(np.median(dev[2:])), gpu_wall_ms=float(np.median(wall[2:])),
                          cpu_ms_all_threads=cpu_all, cpu_threads_all=os.cpu_count(), cpu_ms_1_thread=cpu_1,
                          max_abs_T_diff=float(np.abs(T - want["T"]).max()))))
    g.close()


if __name__ == "__main__":
    main()
