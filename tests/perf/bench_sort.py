#!/usr/bin/env python
"""The radix sort behind every VoxelGrid / index build, measured through the stages that use it (device ms by the
library's CUDA events, results checked bit-exact against the oracle first):
    python tests/perf/bench_sort.py                   # one-sweep sort (default)
    LIOGPU_SORT=3launch python tests/perf/bench_sort.py   # round 1's three launches per pass
One JSON line per stage."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lio_slam_b200 import synth  # noqa: E402
from lio_slam_b200.liogpu import LioGpu  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def med(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    return float(np.median([fn() for _ in range(reps)]))


def main():
    import torch
    o = Oracle("port")
    world = synth.make_world(1234)
    g = LioGpu()
    sort = os.environ.get("LIOGPU_SORT", "onesweep")
    # sweep VoxelGrid: 128-beam sweep at leaf 0.4 (downsampleCurrentScan at config-3 size)
    sweep = synth.to_packed(synth.make_scan(world, synth.path_pose(0.3), 128, seed=5))
    want, _ = o.voxel_grid(sweep, 0.4)
    got, _ = g.voxel_downsample(sweep, 0.4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    d_sweep = torch.from_numpy(sweep).cuda()

    def f1():
        g.voxel_downsample((d_sweep.data_ptr(), sweep.shape[0], 16), 0.4, keep_on_device=True)
        return g.last_gpu_ms()
    l0 = g.launch_count(); f1(); launches = g.launch_count() - l0
    print(json.dumps(dict(stage="sweep VoxelGrid leaf 0.4", sort=sort, n_in=int(sweep.shape[0]), n_out=int(want.shape[0]),
                          device_ms=med(f1), launches=launches)), flush=True)
    # 5 M points (config-4 size): 50 copies of sweeps along the path, leaf 0.5
    clouds = [synth.transform_packed(synth.to_packed(synth.make_scan(world, synth.path_pose(1.0 * k), 64, seed=40 + k)), synth.path_pose(1.0 * k))
              for k in range(44)]
    big = np.concatenate(clouds)[:5_000_000]
    want, _ = o.voxel_grid(big, 0.5)
    got, _ = g.voxel_downsample(big, 0.5)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    d_big = torch.from_numpy(big).cuda()

    def f2():
        g.voxel_downsample((d_big.data_ptr(), big.shape[0], 16), 0.5, keep_on_device=True)
        return g.last_gpu_ms()
    l0 = g.launch_count(); f2(); launches = g.launch_count() - l0
    ms = med(f2)
    print(json.dumps(dict(stage="VoxelGrid 5M leaf 0.5", sort=sort, n_in=int(big.shape[0]), n_out=int(want.shape[0]), device_ms=ms,
                          launches=launches, algorithmic_gbs=(16 * big.shape[0] + 16 * want.shape[0]) / ms / 1e6)), flush=True)
    # 5-NN index build of a 500 k map
    map4 = synth.make_local_map(world, 128, 500_000, 0.2, seed=3, s0=-0.5)
    d_map = torch.from_numpy(map4).cuda()

    def f3():
        g.set_local_map((d_map.data_ptr(), map4.shape[0], 16))
        return g.last_gpu_ms()
    l0 = g.launch_count(); f3(); launches = g.launch_count() - l0
    print(json.dumps(dict(stage="5-NN index build 500k", sort=sort, device_ms=med(f3), launches=launches)), flush=True)
    g.close()


if __name__ == "__main__":
    main()
