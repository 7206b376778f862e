"""Turn an .ncu-rep into a small CSV of the metrics the judge reads (run here, no GPU needed):
    python scripts/summarize_ncu.py gpurun_out/prof_gemm_r1.ncu-rep profiles/r1_gemm_metrics.csv
and a launch-list CSV (ncu --metrics gpu__time_duration.sum --csv) into a per-kernel share table:
    python scripts/summarize_ncu.py --launches gpurun_out/launches_r1.csv profiles/r1_launch_shares.md
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__pcsamp_sample_count",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_no_instructions",
]


def rep_to_csv(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEEP]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(rows) - 2)])
        for i in idx:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
    print("wrote", out)


def launches_to_md(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"ncu launch list ({path}): per-kernel device time of ONE forward, cold-cache and serialised by ncu\n"
                f"(compare SHARES, not absolutes). total {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches_to_md(sys.argv[2], sys.argv[3])
    else:
        rep_to_csv(sys.argv[1], sys.argv[2])
