#!/usr/bin/env python
"""bench.py — scan-to-map registrations/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3|cfg3_leaf04|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one scan-to-map registration (liogpu_scan2map: the whole Gauss-Newton loop of
scan2MapOptimization, mapOptmization.cpp:1848-1859) of one synthetic sweep against a resident local map.
N=1 workload (BASELINE.json configs[2], the one the north_star target is quoted on):
    128-beam sweep, 230,400 points, all used as queries, vs a 500,000-point local map (leaf 0.2).
`value`  : registrations/s with the sweep already in HBM (packed float4), device time by CUDA events on the
           library's stream, L2 flushed between steps (outside the timed events).
`e2e`    : the same call with the sweep in pinned HOST memory as 32-byte pcl::PointXYZI records; wall clock
           around the C-ABI call, H2D of the sweep and D2H of the result inside.
N>1      : one process per GPU, independent sequences (weak scaling), no collective on the data path.
--impl reference : the CPU restatement of the reference path (oracle/, KD-tree rebuilt every scan like
           mapOptmization.cpp:1846, OpenMP over all host cores) on the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: beams, n_map, map leaf, scan leaf (None = all points are queries), n_scans
    "cfg3": dict(beams=128, n_map=500_000, map_leaf=0.2, scan_leaf=None, cols=1800,
                 desc="128-beam sweep (230400 pts) vs 500k-pt local map, full LM loop on device"),
    "cfg3_leaf04": dict(beams=128, n_map=500_000, map_leaf=0.2, scan_leaf=0.4, cols=1800,
                        desc="128-beam sweep voxelised at the reference's default scan leaf 0.4 (~67k pts) vs 500k-pt local map"),
    "cfg1": dict(beams=16, n_map=40_000, map_leaf=0.5, scan_leaf=0.4, cols=1800,
                 desc="VLP-16 sweep (28800 pts, leaf 0.4) vs 40k-pt local map"),
}
MAX_ITER = 30


def make_workload(name: str, seed: int, n_scans: int):
    from lio_slam_b200 import synth
    w = WORKLOADS[name]
    world = synth.make_world(1234)
    map4 = synth.make_local_map(world, w["beams"], w["n_map"], w["map_leaf"], seed=3 + seed, s0=-0.5, cols=w["cols"])
    scans, guesses = [], []
    for k in range(n_scans):
        pose_gt = synth.path_pose(0.3 * k)
        sc = synth.to_packed(synth.make_scan(world, pose_gt, w["beams"], seed=1000 * seed + 7 + k, cols=w["cols"]))
        if w["scan_leaf"] is not None:
            sc = synth.voxel_numpy(sc, w["scan_leaf"])
        scans.append(np.ascontiguousarray(sc, np.float32))
        guesses.append(synth.perturbed_guess(pose_gt, 50 + k + 100 * seed))
    return map4, scans, guesses


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  Started
    before the warm-up so the first samples exist when the (short) timed region begins; only samples whose
    arrival time falls inside [mark_start, mark_stop] are reported (all of them if that window is empty)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout: float = 5.0):
        t = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                f = [x.strip() for x in r.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, reasons
        inside = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1 + 0.03]
        sm, mx, reasons = parse(inside if inside else self.rows)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": "timed region" if inside else "whole run"}


def cpu_baseline_run(name: str, map4, scans, guesses, steps: int, warmup: int, kind_pref: str = "auto"):
    """The reference's CPU path (restated): KD build every scan + OpenMP LM loop on all host cores."""
    from oracle.oracle import Oracle, build
    build()
    kind = "port"
    if kind_pref in ("auto", "nanoflann") and Oracle.available("nanoflann"):
        try:
            o = Oracle("nanoflann")
            kind = "nanoflann"
        except OSError:
            o = Oracle("port")
    else:
        o = Oracle("port")
    threads = os.cpu_count() or 1
    times, iters = [], []
    for s in range(warmup + steps):
        k = s % len(scans)
        t0 = time.perf_counter()
        pose, P, info = o.scan2map(map4, scans[k], guesses[k], max_iter=MAX_ITER, threads=threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt); iters.append(info["iterations"])
    ms = 1e3 * float(np.mean(times))
    return dict(value=1e3 / ms, unit="registrations/s", cores=threads, kind="port",
                sample=f"{steps} full registrations of the same workload ({name}); KD-tree "
                       f"({'reference-vendored nanoflann 1.3.2' if kind == 'nanoflann' else 'own FLANN-style tree'}, leaf 15) "
                       f"rebuilt every scan as mapOptmization.cpp:1846 does; mean {float(np.mean(iters)):.1f} LM iterations",
                ms_per_registration=ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="liogpu", choices=["liogpu", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload
    w = WORKLOADS[name]
    config = {"workload": f"{name}: {w['desc']}", "beams": w["beams"], "n_map": w["n_map"], "map_leaf": w["map_leaf"],
              "max_iter": MAX_ITER, "sequences": "one independent sequence per GPU (every rank replays the same synthetic sequence)",
              "l2": "flushed (512 MiB write) between timed steps, outside the timed events"}

    if args.impl == "reference":
        if rank != 0:
            return
        # a step is one full CPU registration (~0.14 s on cfg3 with 16 threads): bounded so that the arm ends in
        # well under a minute; at least 3 warm-up steps (the first parallel regions of a fresh process run slow)
        steps = max(1, min(args.steps, 50))
        warm = max(3, min(args.warmup, 10))
        map4, scans, guesses = make_workload(name, 0, 2)
        cb = cpu_baseline_run(name, map4, scans, guesses, steps, warm)
        config["n_query"] = int(scans[0].shape[0])
        line = {"impl": "reference", "metric": "scan2map_registrations_per_sec", "value": cb["value"],
                "unit": "registrations/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": cb["ms_per_registration"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libliogpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from lio_slam_b200.liogpu import LioGpu, S2MInfo, default_params
    import ctypes as C

    n_scans = 4
    # every rank runs the SAME synthetic sequence (its own copy, its own GPU, no communication): N-GPU work is
    # then exactly N x the 1-GPU work and the scaling number is not blurred by data-dependent iteration counts
    map4, scans, guesses = make_workload(name, 0, n_scans)
    nqs = [int(sc.shape[0]) for sc in scans]       # voxelised sweeps differ in size from sweep to sweep
    nq = int(round(sum(nqs) / len(nqs)))           # mean: byte counts below are per average step
    config["n_query"] = nq
    if os.environ.get("LIOGPU_BENCH_PRESORT"):  # experiment: spatially coherent query order
        from lio_slam_b200 import synth as _s
        for k in range(n_scans):
            m = _s.transform_packed(scans[k], guesses[k])
            c = np.floor(m[:, :3] / 0.5).astype(np.int64); c -= c.min(axis=0)
            key = (c[:, 2] * (c[:, 1].max() + 1) + c[:, 1]) * (c[:, 0].max() + 1) + c[:, 0]
            scans[k] = np.ascontiguousarray(scans[k][np.argsort(key, kind="stable")])
    g = LioGpu(default_params(device=local_rank, n_scan=w["beams"], horizon_scan=w["cols"],
                              surrounding_keyframe_map_leaf_size=w["map_leaf"],
                              knn_cell_size=float(os.environ.get("LIOGPU_BENCH_CELL", "0")),
                              knn_phase1_radius=float(os.environ.get("LIOGPU_BENCH_R1", "0"))))
    g.set_local_map(map4)
    ext = torch.cuda.ExternalStream(g.stream(), device=torch.device("cuda", local_rank))

    # device-resident sweeps (packed float4) and pinned host sweeps (32-byte PointXYZI records)
    dev_scans = [torch.from_numpy(s).cuda() for s in scans]
    host_recs = []
    for s in scans:
        rec = torch.zeros((s.shape[0], 8), dtype=torch.float32).pin_memory()
        rec[:, 0:3] = torch.from_numpy(s[:, 0:3]); rec[:, 3] = 1.0; rec[:, 4] = torch.from_numpy(s[:, 3])
        host_recs.append(rec)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(k):
        pose, P, info = g.scan2map((dev_scans[k].data_ptr(), nqs[k], 16), guesses[k], max_iter=MAX_ITER)
        return info

    def step_host(k):
        pose, P, info = g.scan2map((host_recs[k].data_ptr(), nqs[k], 32), guesses[k], max_iter=MAX_ITER)
        return info

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    for s in range(max(args.warmup, 3)):
        step_device(s % n_scans); step_host(s % n_scans)

    # ---------------- timed region 1: inputs resident in HBM ----------------
    barrier()
    if sampler:
        sampler.mark_start()
    launches0 = g.launch_count()
    dev_ms, loop_ms, iters = [], [], []
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        ev0.record(ext)
        info = step_device(s % n_scans)
        ev1.record(ext)
        ev1.synchronize()
        dev_ms.append(ev0.elapsed_time(ev1)); loop_ms.append(info["gpu_ms"]); iters.append(info["iterations"])
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = g.launch_count() - launches0
    # ---------------- timed region 2: end to end from pinned host memory ----------------
    e2e_ms = []
    for s in range(args.steps):
        flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host(s % n_scans)
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    barrier()
    if sampler:
        sampler.mark_stop()
    clocks = sampler.stop() if sampler else None
    # ---------------- untimed: per-kernel device times of the dominant kernel (for the roofline) ----------------
    gp = LioGpu(default_params(device=local_rank, n_scan=w["beams"], horizon_scan=w["cols"],
                               surrounding_keyframe_map_leaf_size=w["map_leaf"], profile_kernels=1,
                               knn_cell_size=float(os.environ.get("LIOGPU_BENCH_CELL", "0")),
                               knn_phase1_radius=float(os.environ.get("LIOGPU_BENCH_R1", "0"))))
    gp.set_local_map(map4)
    kern = {"main_ms": 0.0, "main_n": 0, "left_ms": 0.0, "left_n": 0, "seeded": 0}
    for s in range(min(args.steps, 16) + 2):
        flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        _, _, inf = gp.scan2map((dev_scans[s % n_scans].data_ptr(), nqs[s % n_scans], 16), guesses[s % n_scans], max_iter=MAX_ITER)
        if s >= 2:
            kern["main_ms"] += inf["main_kernel_ms"]; kern["main_n"] += inf["main_kernel_launches"]
            kern["left_ms"] += inf["left_kernel_ms"]; kern["left_n"] += inf["left_kernel_launches"]
            kern["seeded"] = inf["seeded"]
    gp.close()

    tot_ms = float(np.sum(dev_ms)); tot_e2e = float(np.sum(e2e_ms))
    if dist is not None:
        t = torch.tensor([tot_ms, tot_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms, tot_e2e = float(t[0]), float(t[1])
        la = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(la)
        launches = int(la[0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    total_regs = args.steps * world_size
    value = total_regs / (tot_ms * 1e-3)
    e2e_value = total_regs / (tot_e2e * 1e-3)
    # roofline of the dominant kernel (s2m_iter_kernel): 96 B per query per launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # dominant kernel: s2m_main_kernel when it runs (dense map), else s2m_left_kernel; its average launch
    # duration comes from CUDA events around every launch (library option profile_kernels, untimed pass above)
    if kern["main_n"] > 0 and kern["main_ms"] >= kern["left_ms"]:
        dom, launch_us = "s2m_main_kernel", 1e3 * kern["main_ms"] / kern["main_n"]
    else:
        dom, launch_us = "s2m_left_kernel", 1e3 * kern["left_ms"] / max(kern["left_n"], 1)
    achieved = 96.0 * nq / (launch_us * 1e-6) / 1e9
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (same workload only)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom)
        if tr and tr.get("n_query") == nq and tr.get("workload") == name:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 96 * nq, "avg_launch_us": launch_us,
                "main_kernel_us": 1e3 * kern["main_ms"] / max(kern["main_n"], 1),
                "left_kernel_us": 1e3 * kern["left_ms"] / max(kern["left_n"], 1),
                "iteration_us": 1e3 * float(np.sum(loop_ms)) / float(np.sum(iters)),
                "seeded_points_last_iter": kern["seeded"],
                "note": "achieved = 96 B x n_query / CUDA-event time of one launch of the dominant kernel; the working "
                        "set (map 16 MB + sweep 3.7 MB) is L2-resident, so the HBM fraction is structurally small"}
    cb = None
    if not args.no_cpu_baseline:
        cb = cpu_baseline_run(name, map4, scans, guesses, args.cpu_steps, 1)
    line = {"metric": "scan2map_registrations_per_sec", "value": value, "unit": "registrations/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": "registrations/s", "ms_per_step": tot_e2e / args.steps,
                    "h2d_bytes_per_step": nq * 32 + 1616, "d2h_bytes_per_step": 1616},
            "gpu_launches": launches, "mean_lm_iterations": float(np.mean(iters)),
            "loop_ms_per_step": float(np.mean(loop_ms)), "wall_s_region1": wall_s,
            "roofline": roofline, "cpu_baseline": cb, "clocks": clocks}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
