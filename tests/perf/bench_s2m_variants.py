#!/usr/bin/env python
"""A/B of the LM-loop implementations on the bench workloads (device-resident sweeps, L2 flushed between steps):
   fused       one persistent cooperative launch per registration, exact no-search certificate on (default)
   fused_nocert  the same launch with the certificate switched off (every point searched every iteration)
   two_kernel  the round-1 path: s2m_main_kernel + s2m_left_kernel per iteration
One JSON line per (workload, variant): mean device ms per registration, per iteration, and the phase split measured
with %globaltimer probes / CUDA events in a separate profiled pass.  Every variant must return the same iteration
counts and nsel history (asserted)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="cfg3,cfg1")
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--variants", default="fused,fused_nocert,two_kernel")
    ap.add_argument("--order", default="sweep", help="sweep (ring-major, as the sensor delivers it) | tileRxC: the sweep is "
                    "re-ordered on the host into tiles of R rings x C columns (R*C = 256 = one thread block) — an experiment on "
                    "the locality of the candidate loads, organised sweeps only")
    args = ap.parse_args()
    import torch
    import bench
    from lio_slam_b200.liogpu import LioGpu, default_params
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    over = {"fused": {"s2m_path": 2}, "fused_nocert": {"s2m_path": 2, "s2m_no_certificate": 1}, "two_kernel": {"s2m_path": 1}}
    for name in args.workloads.split(","):
        w = bench.WORKLOADS[name]
        map4, scans, guesses = bench.make_workload(name, 0, 4)
        if args.order.startswith("tile"):
            R, Cc = (int(x) for x in args.order[4:].split("x"))
            beams, cols = w["beams"], w["cols"]
            if all(sc.shape[0] == beams * cols for sc in scans) and beams % R == 0 and cols % Cc == 0:
                idx = np.arange(beams * cols).reshape(beams // R, R, cols // Cc, Cc).transpose(0, 2, 1, 3).reshape(-1)
                scans = [np.ascontiguousarray(sc[idx]) for sc in scans]
            else:
                print(f"[{name}] not an organised {beams} x {cols} sweep: order left as is", file=sys.stderr)
        dev = [torch.from_numpy(s).cuda() for s in scans]
        ref_hist = None
        for var in args.variants.split(","):
            res = {}
            for prof in (0, 1):
                g = LioGpu(default_params(n_scan=w["beams"], horizon_scan=w["cols"],
                                          surrounding_keyframe_map_leaf_size=w["map_leaf"], profile_kernels=prof, **over[var]))
                g.set_local_map(map4)
                ms, iters, main, left, tail, cert, lo, hist = [], [], 0.0, 0.0, 0.0, [], [], []
                for s in range(args.steps + 3):
                    k = s % len(scans)
                    flush.fill_(s & 0xff)
                    torch.cuda.synchronize()
                    pose, P, info = g.scan2map((dev[k].data_ptr(), scans[k].shape[0], 16), guesses[k])
                    if s < 3:
                        continue
                    ms.append(info["gpu_ms"]); iters.append(info["iterations"])
                    main += info["main_kernel_ms"]; left += info["left_kernel_ms"]; tail += info["tail_ms"]
                    cert.append(info["certified"]); lo.append(info["leftovers"])
                    if prof == 1 and s == 3:
                        res["one_registration"] = dict(
                            main_us=[round(float(x), 1) for x in info["main_us_hist"]],
                            rest_us=[round(float(x), 1) for x in info["rest_us_hist"]],
                            seeded=info["seeded_hist"].tolist(), certified=info["certified_hist"].tolist(),
                            leftovers=info["leftover_hist"].tolist(), nsel=info["nsel_hist"].tolist())
                    if s < 3 + len(scans):
                        hist.append((info["iterations"], info["nsel_hist"].tolist(), pose.tolist()))
                g.close()
                n_it = float(np.sum(iters))
                if prof == 0:
                    res.update(ms_per_registration=float(np.mean(ms)), mean_iterations=float(np.mean(iters)),
                               us_per_iteration=1e3 * float(np.sum(ms)) / n_it)
                else:
                    res.update(main_phase_us_per_iteration=1e3 * main / n_it, rest_us_per_iteration=1e3 * left / n_it,
                               tail_us_per_iteration=1e3 * tail / n_it, certified_last_iter_mean=float(np.mean(cert)),
                               leftovers_last_iter_mean=float(np.mean(lo)), profiled_ms_per_registration=float(np.mean(ms)))
            if ref_hist is None:
                ref_hist = hist
            same = all(a[0] == b[0] and a[1] == b[1] for a, b in zip(hist, ref_hist))
            pose_eq = all(a[2] == b[2] for a, b in zip(hist, ref_hist))
            print(json.dumps(dict(workload=name, variant=var, order=args.order, lib=os.environ.get("LIOGPU_LIB", "default"), n_query=int(scans[0].shape[0]), same_iterations_and_nsel=same,
                                  poses_bit_equal_to_first_variant=pose_eq, **res)), flush=True)
            assert same, "variants disagree on iteration counts / nsel history"


if __name__ == "__main__":
    main()
