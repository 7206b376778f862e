"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over ONE eager forward,
scripts/ncu_forward.py) -> profiles/dram_traffic.json (read by bench.py for `roofline.traffic`) and a markdown table.

    python scripts/dram_traffic.py gpurun_out/r2h/dram_cfg4.csv profiles/r2_dram_launches.csv profiles/r2_dram_traffic.md
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, raw_copy, md = sys.argv[1:4]
rows = collections.defaultdict(lambda: collections.defaultdict(float))   # (id) -> metric -> value
names = {}
for r in csv.DictReader(l for l in open(src) if not l.startswith("==")):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    rows[r["ID"]][r["Metric Name"]] = v * scale
    names[r["ID"]] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("mtt::", "").strip()
agg = collections.OrderedDict()
for i, m in rows.items():
    a = agg.setdefault(names[i], {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0})
    a["launches"] += 1
    a["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
    a["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
    a["time_us"] += m.get("gpu__time_duration.sum", 0.0)
for a in agg.values():
    a["bytes_per_launch"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"]
fam = [a for k, a in agg.items() if k.startswith("gemm")]
gf = {"launches": sum(a["launches"] for a in fam),
      "bytes_per_launch": sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in fam) / max(1, sum(a["launches"] for a in fam)),
      "time_us": sum(a["time_us"] for a in fam)}
out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one "
                 f"eager tp_cfg4 bs 4 parity forward (scripts/ncu_forward.py), {os.path.relpath(raw_copy, ROOT)}",
       "kernels": agg, "gemm_family": gf}
shutil.copyfile(src, raw_copy)
with open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
with open(md, "w") as f:
    f.write("DRAM traffic per kernel of ONE TaskPrompter cfg4 (bs 4, parity) forward, `ncu --metrics dram__bytes_read.sum,\n"
            f"dram__bytes_write.sum,gpu__time_duration.sum --clock-control none` (raw: {os.path.basename(raw_copy)}). "
            "Serialised and cold-cache per launch.\n\n"
            "| kernel | launches | DRAM read MB | DRAM write MB | MB / launch | total us | avg GB/s (of 6582 peak) |\n"
            "|---|---:|---:|---:|---:|---:|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]):
        tot = a["dram_read_bytes"] + a["dram_write_bytes"]
        f.write(f"| {k} | {a['launches']} | {a['dram_read_bytes'] / 1e6:.1f} | {a['dram_write_bytes'] / 1e6:.1f} | "
                f"{a['bytes_per_launch'] / 1e6:.2f} | {a['time_us']:.1f} | {tot / max(a['time_us'], 1e-9) / 1e3:.0f} |\n")
    f.write(f"\nGEMM / conv family: {gf['launches']} launches, {gf['bytes_per_launch'] / 1e6:.1f} MB of DRAM traffic per launch.\n")
print("gemm family:", gf)
