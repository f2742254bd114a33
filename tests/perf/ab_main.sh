#!/bin/bash
# A/B of the main-kernel variants of the two-kernel LM loop (run on the GPU box from the repository root):
#   tests/perf/ab_main.sh TAG "fixed pw wc" "default t128 t64"
# One line per (library build, LIOGPU_MAIN, workload): us per iteration, search + fit phase, the rest, per-iteration detail.
TAG=${1:-ab}; VARS=${2:-"fixed pw"}; LIBS=${3:-"default"}
for l in $LIBS; do
  for v in $VARS; do
    if [ "$l" = default ]; then unset LIOGPU_LIB; else export LIOGPU_LIB=$PWD/lio_slam_b200/libliogpu_$l.so; fi
    LIOGPU_MAIN=$v timeout 200 python tests/perf/bench_s2m_variants.py --variants two_kernel --workloads cfg3,cfg1 \
      > gpurun_out/${TAG}_var_${l}_$v.jsonl 2> gpurun_out/${TAG}_var_${l}_$v.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_var_*.jsonl")):
    for l in open(f):
        d = json.loads(l)
        print(f.split("/")[-1], d["workload"], round(d["us_per_iteration"], 1), round(d["main_phase_us_per_iteration"], 1),
              round(d["rest_us_per_iteration"], 1), d["one_registration"]["main_us"], d["one_registration"]["rest_us"],
              d["poses_bit_equal_to_first_variant"])
PY
