"""CPU: bench.py's reference arm prints one well-formed JSON line (tiny config so it runs in seconds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config",
                          "tp_tiny", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, check=True)
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout     # exactly ONE line on stdout, and it is the JSON
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["value"] > 0 and j["unit"] == "images/s"
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["steps"] == 2 and j["warmup"] == 1 and j["config"]["global_batch"] == 2
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["higher_is_better"] is True


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_bench_line_contract_live(cuda_dev):
    """Runs bench.py itself on the device (tiny configuration, a few steps) and checks the line it prints against the
    driver's contract: every key, value = images / device time, e2e copies declared, launches counted, roofline
    self-consistent, the baseline legs present."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "tp_tiny", "--steps", "4",
                          "--warmup", "3", "--repeats", "3"], capture_output=True, text=True, check=True)
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
              "gpu_eager_baseline", "repeats"):
        assert k in j, k
    assert j["metric"] == "images/sec" and j["unit"] == "images/s" and j["higher_is_better"] is True
    assert j["scaling"] == "weak" and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert "workload" in j["config"] and "model" not in j["config"]
    B = j["config"]["global_batch"]
    assert j["steps"] == 4 and abs(j["value"] - B * 1e3 / j["ms_per_step"]) < 1e-6 * j["value"]
    assert len(j["repeats"]["images_per_s"]) == 3 and min(j["repeats"]["images_per_s"]) <= j["value"] <= max(j["repeats"]["images_per_s"])
    e = j["e2e"]
    assert e["unit"] == j["unit"] and e["value"] > 0
    assert e["h2d_bytes_per_step"] == B * 3 * 64 * 96 * 4 and e["d2h_bytes_per_step"] > 0
    assert j["gpu_launches"] == j["launches_per_step"] * j["steps"] > 0
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = j["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == j["unit"] and c["sample"]
    g = j["gpu_eager_baseline"]
    assert g["fp32"] > 0 and g["tf32"] > 0 and g["bf16_autocast"] > 0
    t = j["train_step"]          # the training step of the same workload on the same ranks (the leg with the collective)
    assert "error" not in t, t
    assert t["metric"] == "train images/sec" and t["value"] > 0 and t["global_batch"] == B
    assert abs(t["value"] - B * 1e3 / t["ms_per_step"]) < 1e-6 * t["value"] and t["last_loss"] == t["last_loss"]
