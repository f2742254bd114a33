#!/usr/bin/env python
"""The online per-sweep pipeline through the C ABI, sweep crossing PCIe once:
liogpu_deskew (kept on device) -> liogpu_downsample_scan2map (resident) -> liogpu_keyframe_put (resident),
for the 16-beam (configs[0]) and 32-beam + IMU (configs[1]) shapes, with the CPU oracle pipeline beside it; then the
two per-keyframe / per-scan extras of the fork: the Scan Context descriptor and publishLocalMap over 30 keyframes.
Prints one JSON line per shape (wall ms per sweep, median of 30)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lio_slam_b200 import synth  # noqa: E402
from lio_slam_b200.liogpu import LioGpu, RESIDENT  # noqa: E402
from oracle.oracle import DeskewParams, Oracle  # noqa: E402


def main():
    o = Oracle("nanoflann") if Oracle.available("nanoflann") else Oracle("port")
    world = synth.make_world(1234)
    threads = os.cpu_count() or 1
    for beams, n_map, name in ((16, 40000, "configs[0] 16-beam"), (32, 40000, "configs[1] 32-beam + 200 Hz IMU deskew")):
        kw = dict(n_scan=beams, downsample_rate=1, point_filter_num=1, lidar_min_front=1.0, lidar_min_back=5.0, lidar_min_left=2.0,
                  lidar_min_right=2.0, lidar_max_range=1000.0, lidar_max_intensity=100.0, mapping_surf_leaf_size=0.4,
                  surrounding_keyframe_map_leaf_size=0.5)
        g = LioGpu(**kw)
        pose_gt = synth.path_pose(0.3)
        scan = synth.make_scan(world, pose_gt, beams, seed=5)
        t0 = 1000.0
        imu = synth.make_imu_table(t0, seed=3)
        map4 = synth.make_local_map(world, beams, n_map, 0.5, seed=3, s0=-0.5)
        guess = synth.perturbed_guess(pose_gt, 8)
        g.set_local_map(map4)
        dp = DeskewParams(beams, 1, 1, 1.0, 5.0, 2.0, 2.0, 1000.0, 100.0)
        # 30 earlier keyframes (utility.h:219 localMapKeyFramesNumber) for the per-scan publishLocalMap
        kf_poses = []
        for k in range(30):
            pk = synth.path_pose(0.3 - 1.0 * (k + 1))
            dsk_k, _ = o.voxel_grid(synth.to_packed(synth.make_scan(world, pk, beams, seed=200 + k, cols=900)), 0.4)
            g.keyframe_put(k, dsk_k)
            kf_poses.append(pk.astype(np.float32))
        kf_ids = list(range(30))
        kf_poses = np.array(kf_poses, np.float32)
        stage = {"deskew": [], "downsample+scan2map": [], "keyframe_put": [], "total": [], "scancontext": [],
                 "publishLocalMap": [], "total_with_scancontext_and_local_map": []}
        for s in range(35):
            a = time.perf_counter()
            n_dsk, _ = g.deskew(scan, t0, *imu, True, keep_on_device=True)
            b = time.perf_counter()
            pose, P, info = g.downsample_scan2map(RESIDENT, guess, keep_ds_on_device=True)
            c = time.perf_counter()
            g.keyframe_put(1000 + (s % 4), RESIDENT)
            d = time.perf_counter()
            g.make_scancontext(RESIDENT)                      # SCInputType::SINGLE_SCAN_FEAT (mapOptmization.cpp:2158-2160)
            e = time.perf_counter()
            lm, lm_info, _ = g.publish_local_map(kf_ids, kf_poses, pose, use_removing_outliers=0,
                                                 local_mapping_surf_leaf_size=0.2)      # jeep.yaml settings
            f = time.perf_counter()
            if s >= 5:
                stage["deskew"].append(b - a); stage["downsample+scan2map"].append(c - b)
                stage["keyframe_put"].append(d - c); stage["total"].append(d - a)
                stage["scancontext"].append(e - d); stage["publishLocalMap"].append(f - e)
                stage["total_with_scancontext_and_local_map"].append(f - a)
        cpu = []
        h = None
        for s in range(4):
            a = time.perf_counter()
            dsk = o.deskew(scan, dp, t0, *imu, True)
            ds, _ = o.voxel_grid(dsk, 0.4)
            ref_pose, _, ref_info = o.scan2map(map4, ds, guess, threads=threads)
            cpu.append(time.perf_counter() - a)
        assert info["iterations"] == ref_info["iterations"] and np.abs(pose - ref_pose).max() <= 1e-4
        print(json.dumps({"shape": name, "n_raw": int(scan.shape[0]), "n_deskewed": int(n_dsk), "n_ds": info["n_ds"],
                          "iterations": info["iterations"],
                          "gpu_wall_ms": {k: round(1e3 * float(np.median(v)), 4) for k, v in stage.items()},
                          "cpu_ms": round(1e3 * float(np.median(cpu[1:])), 3), "cpu_threads": threads,
                          "local_map_points": [lm_info["n_concat"], lm_info["n_out"]],
                          "h2d_bytes_per_sweep": int(scan.shape[0]) * 32, "pose_bit_equal": bool(np.array_equal(pose, ref_pose))}))
        g.close()


if __name__ == "__main__":
    main()
