#!/usr/bin/env python
"""Batch offline mapping on ONE GPU (BASELINE configs[4] has more sequences than GPUs): K independent
sequences, each with its own liogpu context + stream + host thread, registered concurrently.  A single
registration is latency bound and leaves most SMs idle, so several sequences per GPU raise the aggregate
registrations/s.  ctypes releases the GIL during the C call, so plain Python threads are enough here.
Prints one JSON line per K."""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lio_slam_b200.liogpu import LioGpu, default_params  # noqa: E402


def main():
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    w = bench.WORKLOADS[name]
    map4, scans, guesses = bench.make_workload(name, 0, 4)
    nq = scans[0].shape[0]
    dev_scans = [torch.from_numpy(s).cuda() for s in scans]
    steps = 200 if name == "cfg1" else 60
    for K in (1, 2, 4, 8):
        ctxs = []
        for _ in range(K):
            g = LioGpu(default_params(n_scan=w["beams"], horizon_scan=w["cols"], surrounding_keyframe_map_leaf_size=w["map_leaf"]))
            g.set_local_map(map4)
            ctxs.append(g)
        def worker(g, out, k):
            for s in range(5):
                g.scan2map((dev_scans[s % 4].data_ptr(), nq, 16), guesses[s % 4])
            barrier.wait()
            t0 = time.perf_counter()
            for s in range(steps):
                g.scan2map((dev_scans[(s + k) % 4].data_ptr(), nq, 16), guesses[(s + k) % 4])
            out[k] = time.perf_counter() - t0
        barrier = threading.Barrier(K)
        out = [0.0] * K
        th = [threading.Thread(target=worker, args=(ctxs[k], out, k)) for k in range(K)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        wall = max(out)
        print(json.dumps({"workload": name, "contexts_on_one_gpu": K, "registrations_per_s": K * steps / wall,
                          "ms_per_registration_per_sequence": 1e3 * wall / steps}))
        for g in ctxs:
            g.close()


if __name__ == "__main__":
    main()
