"""The bench records committed under profiles/ (what profiles/README.md and DESIGN.md quote) are single JSON lines that
keep the driver's contract: they parse, carry the contract keys, name the workload of BASELINE.json's metric, and their
derived figures are consistent with each other."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def load(name):
    lines = [l for l in open(os.path.join(PROFILES, name)).read().splitlines() if l.startswith("{")]
    assert len(lines) == 1, name
    return json.loads(lines[0])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_round2_gpu_arm_records(n):
    d = load(f"r02_bench_n{n}.json")
    assert d["metric"] == "scan2map_registrations_per_sec" and d["unit"] == "registrations/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n and d["steps"] >= 20 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["data"] == "synthetic" and d["vs_baseline"] is None and d["config"]["workload"].startswith("cfg3")
    assert abs(d["value"] - n * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]           # whole-job aggregate over the N ranks
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 7_000_000 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["kernel"] == "s2m_main_kernel" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_us"] * 1e-6) / 1e9) < 1e-6 * r["achieved"]
    assert r["algorithmic_bytes_per_launch"] == 96 * d["config"]["n_query"]
    assert d["gpu_launches"] > 0 and not d["clocks"]["reasons"] and d["clocks"]["sm_mhz"] >= 0.95 * d["clocks"]["sm_max_mhz"]
    c4, c5 = d["cfg4"], d["cfg5"]
    assert c4["bit_equal_to_1gpu"] is True and c4["n_gpus"] == n and c4["n_points"] > 4_000_000
    assert c5["n_gpus"] == n and c5["scaling"] == "strong" and c5["total_scans"] == 8 * c5["scans_per_sequence"]
    assert len(c5["wall_s_repetitions"]) >= 3 and max(c5["wall_s_repetitions"]) < 1.1 * min(c5["wall_s_repetitions"])
    assert abs(c5["value"] - c5["total_scans"] / c5["wall_s"]) < 1e-6 * c5["value"]
    assert c5["max_position_error_vs_ground_truth_m"] < 0.15
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and set(cb["by_number_of_cores"]) >= {"4", "12"}
        assert cb["value_kd_build_excluded"] > cb["value"] > 0
        like = d["like_for_like"]
        assert like["ratio_both_rebuild_every_scan"] > like["ratio_neither_rebuilds"] > 20      # north_star: >= 20x
        assert 0 < like["keyframes_per_scan"] < 1
        assert d["lm_loop_ab"]["fused_one_launch"]["poses_bit_equal_to_two_kernel"] is True
        assert d["cfg1"]["value"] > 0 and d["cfg1"]["cpu_baseline"]["value"] > 0
    else:
        assert c5["weak_companion"]["value"] > c5["value"]


def test_round2_reference_arm_record():
    d = load("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "scan2map_registrations_per_sec" and d["config"]["workload"].startswith("cfg3")
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]
    g = load("r02_bench_n1.json")
    assert g["config"]["workload"] == d["config"]["workload"]           # both arms on the same workload
    assert g["e2e"]["value"] / d["value"] > 20                            # the headline ratio the driver computes
