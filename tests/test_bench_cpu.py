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
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["higher_is_better"] is True
