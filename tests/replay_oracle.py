"""The per-scan path of liorf_replay.cpp (cloudHandler -> laserCloudInfoHandler, imageProjection.cpp:206 /
mapOptmization.cpp:432-506) driven with the CPU oracle: deskew -> extractNearby -> extractCloud -> downsampleCurrentScan
-> scan2MapOptimization -> keyframe gate.  Checker for the host mirror + CUDA library (tests) and the CPU arm of the
batch-mapping bench; never imported by the product."""
import time

import numpy as np

from lio_slam_b200 import synth


def save_frame(last_kf, pose, dist_th=1.0, ang_th=0.2):
    """saveFrame (mapOptmization.cpp:1909-1928): relative motion since the last keyframe, f64 like the host mirror."""
    if last_kf is None:
        return True
    Ra, Rb = synth.rpy_to_R(*last_kf[:3].astype(float)), synth.rpy_to_R(*pose[:3].astype(float))
    D = Ra.T @ Rb
    loc = Ra.T @ (pose[3:].astype(float) - last_kf[3:].astype(float))
    roll, pitch, yaw = np.arctan2(D[2, 1], D[2, 2]), np.arcsin(-D[2, 0]), np.arctan2(D[1, 0], D[0, 0])
    return not (abs(roll) < ang_th and abs(pitch) < ang_th and abs(yaw) < ang_th and np.linalg.norm(loc) < dist_th)


def sweep_records(seq, s):
    raw = seq["raw"]
    a = raw.numpy() if hasattr(raw, "numpy") else raw
    lo, hi = int(seq["offs"][s]), int(seq["offs"][s + 1])
    return a.reshape(-1, 8)[lo:hi].view(np.uint8).reshape(-1, 32).view(synth.XYZIRT_DTYPE).reshape(-1)


def replay_sequence_oracle(o, seq, first=0, count=None, n_scan=64, downsample_rate=2, point_filter_num=5, scan_leaf=0.4,
                           map_leaf=0.5, radius=50.0, density=2.0, threads=8, min_range=1.0, max_range=1000.0):
    from oracle.oracle import DeskewParams
    dp = DeskewParams(n_scan, downsample_rate, point_filter_num, min_range, min_range, min_range, min_range, max_range, 100.0)
    n_all = len(seq["offs"]) - 1
    n = n_all - first if count is None else count
    key_poses, key_times, kf = [], [], []
    matP, deg = np.zeros((6, 6), np.float32), 0
    cur_sel, map4 = None, None
    poses, iters, nds = np.zeros((n, 6), np.float32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    stats = dict(scans=0, registered=0, keyframes=0, map_rebuilds=0, lm_iterations=0, wall_ms=0.0, kd_build_ms=0.0, loop_ms=0.0)
    t_all = time.perf_counter()
    for k in range(n):
        s = first + k
        t = float(seq["times"][s])
        imu = seq["imu"][s]
        cloud = o.deskew(sweep_records(seq, s), dp, t, imu[0], imu[1], imu[2], imu[3], True)
        pose = np.array(seq["guesses"][s], np.float32)
        if key_poses:
            key3d = np.array([[p[3], p[4], p[5], float(j)] for j, p in enumerate(key_poses)], np.float32)
            ids = o.extract_nearby(key3d, np.array(key_times), t, radius, density)
            sel = (tuple(ids.tolist()), tuple(np.concatenate([key_poses[j] for j in ids]).tolist()) if len(ids) else ())
            if sel != cur_sel:
                map4, _ = o.build_local_map([kf[j] for j in ids], np.array([key_poses[j] for j in ids], np.float32), map_leaf,
                                            threads=threads)
                cur_sel = sel
                stats["map_rebuilds"] += 1
        ds, _ = o.voxel_grid(cloud, scan_leaf)
        if key_poses:
            pose, matP, info = o.scan2map(map4, ds, pose, matP=matP, degenerate=deg, threads=threads)
            deg = info["is_degenerate"]
            iters[k] = info["iterations"]
            stats["registered"] += 1
            stats["lm_iterations"] += int(info["iterations"])
            stats["kd_build_ms"] += info["ms_build"]; stats["loop_ms"] += info["ms_loop"]
        if save_frame(key_poses[-1] if key_poses else None, pose):
            kf.append(ds); key_poses.append(pose.copy()); key_times.append(t)
        poses[k] = pose
        nds[k] = ds.shape[0]
        stats["scans"] += 1
    stats["wall_ms"] = 1e3 * (time.perf_counter() - t_all)
    stats["keyframes"] = len(key_poses)
    return poses, iters, nds, stats
