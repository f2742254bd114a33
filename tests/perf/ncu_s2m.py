"""The command profiled by ncu for profiles/r02_ncu_full_s2m_*.csv: the config-3 registration (128-beam sweep, 230,400
points vs a 500,000-point map) with the two-kernel loop and then with the one-launch loop.

    ncu --set full --clock-control none --import-source on -k regex:"s2m_main_kernel|s2m_left_kernel|s2m_fused_kernel" \\
        -s 6 -c 14 -o gpurun_out/r02_s2m python tests/perf/ncu_s2m.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402
from lio_slam_b200.liogpu import LioGpu, default_params  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = bench.WORKLOADS[name]
map4, scans, guesses = bench.make_workload(name, 0, 2)
dev = [torch.from_numpy(s).cuda() for s in scans]
for path in (1, 2):
    g = LioGpu(default_params(n_scan=w["beams"], horizon_scan=w["cols"], surrounding_keyframe_map_leaf_size=w["map_leaf"],
                              s2m_path=path))
    g.set_local_map(map4)
    for k in range(2):
        pose, P, info = g.scan2map((dev[k].data_ptr(), scans[k].shape[0], 16), guesses[k])
        print(path, k, info["iterations"], info["gpu_ms"], info["kernel_launches"])
    g.close()
