"""CPU: the bench line committed from the last B200 run of this build (profiles/r1s_bench_parity.json, written by
`python bench.py --gpus 1 --steps 20 --warmup 3`) carries every key the driver's contract names, with consistent values.
Guards the JSON contract against accidental edits of bench.py's line assembly between GPU runs."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [l for l in f.read().splitlines() if l.strip()]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_committed_bench_line_follows_the_contract():
    j = _line("r1s_bench_parity.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in j, k
    assert j["metric"] == "images/sec" and j["unit"] == "images/s" and j["higher_is_better"] is True
    assert j["scaling"] == "weak" and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert "workload" in j["config"] and "model" not in j["config"]
    B = j["config"]["global_batch"]
    assert abs(j["value"] - B * 1e3 / j["ms_per_step"]) < 1e-6 * j["value"]          # value = images / device time
    e = j["e2e"]
    assert e["unit"] == j["unit"] and 0 < e["value"] <= j["value"]
    assert e["h2d_bytes_per_step"] == B * 3 * 512 * 512 * 4 and e["d2h_bytes_per_step"] > 0
    assert j["gpu_launches"] == j["launches_per_step"] * j["steps"] > 0
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = j["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == j["unit"] and c["sample"]
    k = j["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not (set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                      "sw_thermal_slowdown"})


def test_committed_reference_arm_line():
    j = _line("r1s_bench_reference_arm.json")
    assert j["impl"] == "reference" and j["metric"] == "images/sec" and j["unit"] == "images/s"
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["kind"] == "port"
