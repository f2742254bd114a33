#!/usr/bin/env python
"""Every entry point of the CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer (the reference has
no sanitizer runs; the checker should at least be clean itself).  Usage (CPU only, ~1 min):

    mkdir -p /tmp/asan && g++ -O1 -g -std=c++14 -fPIC -shared -fopenmp -ffp-contract=off -fsanitize=address,undefined \
        -fno-omit-frame-pointer -o /tmp/asan/liboracle.so oracle/liorf_oracle.cpp
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
        LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) python tests/perf/oracle_sanitize.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lio_slam_b200 import synth
import oracle.oracle as om
om._HERE = os.environ.get('ORACLE_SANITIZED_DIR', '/tmp/asan')   # load the sanitized build
o = om.Oracle("port")
world = synth.make_world(1234)
pose = synth.path_pose(0.0)
scan = synth.make_scan(world, pose, 16, seed=11, cols=450)
s4 = synth.to_packed(scan)
ds, ov = o.voxel_grid(s4, 0.4); print("voxel", ds.shape, ov)
ds2, ov2 = o.voxel_grid(s4, 0.001); print("guard", ds2.shape, ov2)
map4 = synth.make_local_map(world, 16, 8000, 0.5, seed=5, s0=-0.5, cols=450, max_poses=16)
guess = synth.perturbed_guess(pose, 3)
p, P, info = o.scan2map(map4, ds, guess, threads=4); print("s2m", info["iterations"], info["converged"])
r = o.surf_optimization(map4, ds, pose6=guess, threads=4); print("surf", r["flag"].sum())
clouds = [ds, map4[:3000]]; poses = np.array([pose, np.zeros(6)], np.float32)
lm, li, md = o.publish_local_map(clouds, poses, pose.astype(np.float32), leaf=0.3, threads=4); print("lmap", lm.shape, li["n_after_sor"])
lm, li, md = o.publish_local_map(clouds, poses, pose.astype(np.float32), leaf=0.3, brute=True, threads=4); print("lmap brute", lm.shape)
src = o.transform_cloud(ds, synth.perturbed_guess(pose, 5, trans=(0.4, 0.3, 0.05)))
ic = o.icp_align(src, map4, threads=4); print("icp", ic["iterations"], ic["state"])
ic = o.icp_align(src + np.float32([500, 0, 0, 0]), map4, max_correspondence_distance=5.0, threads=4); print("icp none", ic["state"])
sc = o.make_scancontext(s4); print("sc", (sc[0] != 0).sum())
n = 300
xyz = np.array([synth.path_pose(0.7 * k)[3:6] for k in range(n)])
key3d = np.c_[xyz, np.arange(n)].astype(np.float32)
print("nearby", o.extract_nearby(key3d, 0.4 * np.arange(n), 0.4 * n).shape)
from oracle.oracle import DeskewParams
imu = synth.make_imu_table(1000.0, seed=3)
d = o.deskew(scan, DeskewParams(16, 1, 1, 1.0, 5.0, 2.0, 2.0, 1000.0, 100.0), 1000.0, *imu, True); print("deskew", d.shape)
bl, _ = o.build_local_map(clouds, poses, 0.5, threads=4); print("build", bl.shape)
print("ALL OK")
