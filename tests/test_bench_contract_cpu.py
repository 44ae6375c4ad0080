"""The bench.py JSON-line contract, checked on the committed lines of both arms (profiles/): every key the driver and
the judge read is present, typed and self-consistent.  (bench.py itself needs a GPU; `--impl cpu` is too slow for the
CPU suite.)"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


BASE = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float,
        "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict, "clocks": dict}


@pytest.mark.parametrize("name", ["r01_bench_ours_final.json", "r01_bench_reference_final.json",
                                  "r01_bench_ours_final_2gpu.json", "r02_bench_ours_final.json",
                                  "r02_bench_reference_final.json", "r02_bench_ours_8gpu_midround.json"])
def test_bench_line_has_the_contract_keys(name):
    d = _line(name)
    for k, t in BASE.items():
        assert k in d and isinstance(d[k], t), (k, type(d.get(k)))
    assert "vs_baseline" in d and d["vs_baseline"] is None            # BASELINE.md publishes no number for this metric
    assert d["metric"].startswith("fwd+bwd renders/sec") and d["unit"] == "renders/s"
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "1200x680" in d["config"]["workload"] and d["config"]["gaussians"] == 1_000_000
    assert abs(d["value"] - d["n_gpus"] * 1000.0 / d["ms_per_step"]) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 65_792_000 and e["d2h_bytes_per_step"] == 13_056_000
    assert e["value"] > 0 and e["value"] != d["value"]                 # measured separately, not a copy of `value`
    c = d["clocks"]
    assert set(["sm_mhz", "sm_max_mhz", "reasons"]) <= set(c)
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if c["sm_mhz"] is not None:
        assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert "cpu_baseline" in d and {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) or d["n_gpus"] > 1


def test_ours_line_counts_its_own_kernels_and_reference_line_is_tagged():
    ours, ref = _line("r01_bench_ours_final.json"), _line("r01_bench_reference_final.json")
    assert ours["gpu_launches"] == 8 * ours["steps"] and "impl" not in ours
    assert ref["impl"] == "reference" and ref["gpu_launches"] == 0
    assert ours["config"]["workload"] == ref["config"]["workload"] and ours["config"]["num_rendered"] == ref["config"]["num_rendered"]
    assert ours["roofline"]["kernel"] == "blend_backward" and ours["roofline"]["traffic"] > ours["roofline"]["algorithmic_bytes_per_launch"]
    assert ours["cpu_baseline"]["kind"] == "port" and ours["cpu_baseline"]["cores"] >= 1
    # the headline claims of DESIGN.md section 5
    assert ours["value"] / ref["value"] > 2.0 and ours["e2e"]["value"] / ref["e2e"]["value"] > 2.0
    assert ours["mapping"]["value"] / ref["mapping"]["value"] > 6.0


def test_round2_lines():
    """Round-2 additions: own-kernel launch count incl. the radix sort, the eager drop-in number, traffic read from the
    committed ncu capture, the mapping line on the view-filling map with the stock reference loop."""
    ours, ref = _line("r02_bench_ours_final.json"), _line("r02_bench_reference_final.json")
    assert ours["gpu_launches"] == 16 * ours["steps"] and ref["impl"] == "reference"
    assert ours["config"]["num_rendered"] == ref["config"]["num_rendered"] == 3097789
    assert ours["value"] / ref["value"] > 3.0 and ours["e2e"]["value"] / ref["e2e"]["value"] > 2.5
    assert ours["eager_drop_in"]["value"] / ref["value"] > 2.0
    assert len(ours["e2e"]["runs_ms_per_step"]) == 3
    rf = ours["roofline"]
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["ours"][rf["kernel"]]["dram_bytes"]
    assert rf["kernel"] == "blend_backward" and rf["traffic"] == traffic > rf["algorithmic_bytes_per_launch"]
    mo, mr = ours["mapping"], ref["mapping"]
    assert mo["workload"] == mr["workload"] == "view_filling_1000000" and mo["steps"] >= 50
    assert mo["num_rendered"] > 1.5 * mo["visible"] and mo["visible"] > 0.99 * mo["gaussians"]
    assert "stock get_loss" in mr["impl_note"] and mo["value"] / mr["value"] > 6.0
    m8 = _line("r02_bench_ours_8gpu_midround.json")["mapping"]
    assert m8["keyframes_per_step"] == 8 and m8["allreduce_ms"] > 0 and m8["value"] > 6.0 * 456.0
