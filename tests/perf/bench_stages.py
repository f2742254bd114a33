#!/usr/bin/env python
"""Per-stage timings of the path outside the LM loop: deskew (config 2), sweep VoxelGrid,
local-map rebuild = transform + concatenate + VoxelGrid + grid index (config 4 sizes), publishLocalMap (§8 f2),
index build alone.
GPU device time (CUDA events inside the library) and wall time through the C ABI from host buffers, with the
CPU oracle timed beside each stage.  Prints one JSON line per stage; results are checked bit-exact first."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lio_slam_b200 import sharding, synth  # noqa: E402
from lio_slam_b200.liogpu import LioGpu  # noqa: E402
from oracle.oracle import DeskewParams, Oracle  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), out


def main():
    o = Oracle("port")
    world = synth.make_world(1234)
    res = []
    # ---- config 2: 32-beam sweep, 200 Hz IMU deskew
    g = LioGpu(n_scan=32, downsample_rate=1, point_filter_num=1, lidar_min_front=2.0, lidar_min_back=10.0,
               lidar_min_left=2.0, lidar_min_right=2.0, lidar_max_range=100.0, lidar_max_intensity=100.0)
    scan32 = synth.make_scan(world, synth.path_pose(0.5), 32, seed=91)
    t0 = 1700000100.0
    imu = synth.make_imu_table(t0, seed=12)
    dp = DeskewParams(32, 1, 1, 2.0, 10.0, 2.0, 2.0, 100.0, 100.0)
    cpu_ms, want = timeit(lambda: o.deskew(scan32, dp, t0, *imu, True))
    wall_ms, got = timeit(lambda: g.deskew(scan32, t0, *imu, True)[0])
    dev_ms = g.last_gpu_ms()
    assert np.abs(got - want).max() <= 1e-5
    res.append(dict(stage="deskew (config 2)", n_in=int(scan32.shape[0]), n_out=int(want.shape[0]), gpu_device_ms=dev_ms,
                    gpu_wall_ms=wall_ms, cpu_ms=cpu_ms, cpu_threads=1, bit_equal=bool(np.array_equal(got, want)),
                    algorithmic_bytes=32 * int(scan32.shape[0]) + 16 * int(want.shape[0])))
    g.close()
    # ---- sweep VoxelGrid: 128-beam sweep, leaf 0.4
    g = LioGpu()
    scan128 = synth.to_packed(synth.make_scan(world, synth.path_pose(0.0), 128, seed=7))
    cpu_ms, (want, _) = timeit(lambda: o.voxel_grid(scan128, 0.4), reps=3, warm=1)
    wall_ms, (got, _) = timeit(lambda: g.voxel_downsample(scan128, 0.4))
    dev_ms = g.last_gpu_ms()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    res.append(dict(stage="VoxelGrid of a 128-beam sweep, leaf 0.4", n_in=int(scan128.shape[0]), n_out=int(want.shape[0]),
                    gpu_device_ms=dev_ms, gpu_wall_ms=wall_ms, cpu_ms=cpu_ms, cpu_threads=1, bit_equal=True,
                    algorithmic_bytes=16 * int(scan128.shape[0]) + 16 * int(want.shape[0])))
    # ---- config 4: local-map rebuild from 50 keyframes (~100k points each)
    clouds, poses = [], []
    for k in range(50):
        p = synth.path_pose(-0.5 * k)
        sc = synth.to_packed(synth.make_scan(world, p, 64, seed=700 + k))
        clouds.append(sc[: 100000]); poses.append(p.astype(np.float32))
    poses = np.array(poses)
    total = int(sum(c.shape[0] for c in clouds))
    cpu_ms, (want, _) = timeit(lambda: o.build_local_map(clouds, poses, 0.5, threads=os.cpu_count()), reps=2, warm=1)
    for k, c in enumerate(clouds):
        g.keyframe_put(k, c)
    ids = list(range(50))
    wall_ms, (n_map, _) = timeit(lambda: g.build_local_map(ids, poses, 0.5, fetch=False))
    dev_ms = g.last_gpu_ms()
    got, _ = g.build_local_map(ids, poses, 0.5)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    res.append(dict(stage="local-map rebuild (config 4): transform + concat + VoxelGrid leaf 0.5 + grid index, 50 keyframes",
                    n_in=total, n_out=int(want.shape[0]), gpu_device_ms=dev_ms, gpu_wall_ms=wall_ms, cpu_ms=cpu_ms,
                    cpu_threads=os.cpu_count(), bit_equal=True, algorithmic_bytes=16 * total + 16 * int(want.shape[0])))
    # the same rebuild tiled 2/4/8 ways on this one GPU (what each of N GPUs would do; outputs concatenated)
    raw = np.concatenate([o.transform_cloud(c, p) for c, p in zip(clouds, poses)])
    for tiles in (2, 4, 8):
        tile, _ = sharding.plan_voxel_tiles(raw, 0.5, tiles)
        per = []
        outs = []
        for t in range(tiles):
            pts = sharding.shard_points(raw, tile, t)
            ms, (out, _) = timeit(lambda: g.voxel_downsample(pts, 0.5), reps=3, warm=1)
            per.append(g.last_gpu_ms()); outs.append(out)
        cat = np.concatenate(outs)
        assert np.array_equal(cat.view(np.uint32), want.view(np.uint32))
        res.append(dict(stage=f"config 4 tiled {tiles} ways: VoxelGrid per tile (device ms, max over tiles = N-GPU critical path)",
                        n_in=total, tiles=tiles, gpu_device_ms_max=max(per), gpu_device_ms_sum=sum(per), bit_equal=True))
    # ---- publishLocalMap (every scan): 50 keyframes (6t.yaml) of voxelised 32-beam sweeps
    g.keyframe_clear()
    kf, kposes = [], []
    for k in range(50):
        p = synth.path_pose(0.5 * k)
        ds, _ = o.voxel_grid(synth.to_packed(synth.make_scan(world, p, 32, seed=700 + k, cols=900)), 0.4)
        kf.append(ds); kposes.append(p.astype(np.float32))
    kposes = np.array(kposes)
    for k, c in enumerate(kf):
        g.keyframe_put(k, c)
    now = kposes[-1]
    for label, okw, gkw in (
            ("utility.h defaults: outlier filter meanK 10 + leaf 0.01 (overflow guard)", dict(), dict()),
            ("outlier filter meanK 10 + VoxelGrid leaf 0.2", dict(leaf=0.2), dict(local_mapping_surf_leaf_size=0.2)),
            ("jeep.yaml: no outlier filter, VoxelGrid leaf 0.2", dict(leaf=0.2, use_removing_outliers=False),
             dict(local_mapping_surf_leaf_size=0.2, use_removing_outliers=0))):
        cpu1_ms, (want, winfo, _) = timeit(lambda: o.publish_local_map(kf, kposes, now, threads=1, **okw), reps=2, warm=1)
        cpuN_ms, _ = timeit(lambda: o.publish_local_map(kf, kposes, now, threads=os.cpu_count(), **okw), reps=2, warm=1)
        wall_ms, (got, info, _) = timeit(lambda: g.publish_local_map(ids, kposes, now, **gkw))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        res.append(dict(stage="publishLocalMap, 50 keyframes: " + label, n_in=info["n_concat"], n_cropped=info["n_cropped"],
                        n_after_sor=info["n_after_sor"], n_out=info["n_out"], gpu_device_ms=info["gpu_ms"],
                        gpu_wall_ms=wall_ms, cpu_ms=cpu1_ms, cpu_threads=1, cpu_ms_all_threads=cpuN_ms,
                        cpu_threads_all=os.cpu_count(), bit_equal=True,
                        algorithmic_bytes=16 * info["n_concat"] + 16 * info["n_out"]))
    # ---- extractNearby: key-pose selection over a long run
    rng = np.random.default_rng(12)
    for n_key in (1000, 10000):
        ang = np.cumsum(rng.normal(0, 0.05, n_key))
        xyz = np.cumsum(np.c_[np.cos(ang), np.sin(ang), np.zeros(n_key)] * 0.8, axis=0)
        key3d = np.c_[xyz, np.arange(n_key)].astype(np.float32)
        kt = 0.4 * np.arange(n_key)
        cpu_ms, want_ids = timeit(lambda: o.extract_nearby(key3d, kt, kt[-1] + 0.05), reps=3, warm=1)
        wall_ms, (got_ids, _) = timeit(lambda: g.extract_nearby(key3d, kt, kt[-1] + 0.05))
        assert np.array_equal(got_ids, want_ids)
        res.append(dict(stage=f"extractNearby: radius search + density filter + snap + recency, {n_key} key poses", n_in=n_key,
                        n_out=int(want_ids.shape[0]), gpu_device_ms=g.last_gpu_ms(), gpu_wall_ms=wall_ms, cpu_ms=cpu_ms,
                        cpu_threads=1, bit_equal=True))
    # ---- index build alone (kdtreeSurfFromMap->setInputCloud), 500k-point map
    map4 = synth.make_local_map(world, 128, 500000, 0.2, seed=3, s0=-0.5)
    cpu_ms, h = timeit(lambda: o.index_build(map4), reps=3, warm=1)
    wall_ms, _ = timeit(lambda: g.set_local_map(map4))
    dev_ms = g.last_gpu_ms()
    res.append(dict(stage="5-NN index build, 500k-point map (KD-tree on CPU, sorted grid on GPU)", n_in=500000,
                    gpu_device_ms=dev_ms, gpu_wall_ms=wall_ms, cpu_ms=cpu_ms, cpu_threads=1, algorithmic_bytes=40 * 500000))
    g.close()
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
