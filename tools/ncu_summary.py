"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

  python tools/ncu_summary.py [round_tag] [--rep] [raw_csv ...]

Inputs : gpurun_out/launches_<tag>.csv  (ncu --metrics gpu__time_duration.sum launch list)
         gpurun_out/prof_<tag>.ncu-rep  (ncu --set full capture), and / or raw CSV exports of such captures
         (`ncu -i x.ncu-rep --page raw --csv`, made on the GPU box when the report is too large to bring back);
         kernels found in a later input replace those of an earlier one, kernels only in the existing
         profiles/ncu_<tag>_summary.json are kept.
Outputs: profiles/launches_<tag>.csv (copy), profiles/ncu_<tag>_summary.json, profiles/ncu_<tag>_summary.md
bench.py reads profiles/ncu_<tag>_summary.json for roofline.traffic (DRAM bytes per launch of the dominant kernel).
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
summary = {"tag": tag, "launches": {}, "kernels": {}}
prev_json = os.path.join(out_dir, f"ncu_{tag}_summary.json")
if os.path.exists(prev_json):
    summary = json.load(open(prev_json))
extra_csv = [a for a in sys.argv[2:] if a != "--rep"]
use_rep = "--rep" in sys.argv[2:]  # ingest gpurun_out/prof_<tag>.ncu-rep (off by default: a stale report must not win)

launch_csv = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(launch_csv):
    shutil.copy(launch_csv, os.path.join(out_dir, f"launches_{tag}.csv"))
    rows = list(csv.reader(l for l in open(launch_csv) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[ix["Kernel Name"]]].append(float(r[ix["Metric Value"]]))
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        summary["launches"][k] = {"n": len(v), "mean_us": sum(v) / len(v) / 1e3, "share": sum(v) / tot}

rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active": "pipe_tc_pct",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "pipe_tmem_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tc_cycles_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_hmma_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_cycles_pct",
    "sm__icc_request_hit_rate.pct": "icache_hit_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
def ingest(raw_text):
    rows = list(csv.reader(raw_text.splitlines()))
    hdr, units = rows[0], rows[1]
    fresh = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        k = {}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                k[WANT[h]] = v * UNIT.get(units[i], 1.0) if WANT[h] in ("duration", "dram_read", "dram_write") else v
        k["dram_bytes"] = k.get("dram_read", 0.0) + k.get("dram_write", 0.0)
        fresh.setdefault(name, []).append(k)
    summary["kernels"].update(fresh)


if os.path.exists(rep) and use_rep:
    ingest(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
for path in extra_csv:
    ingest(open(path).read())

try:  # the build the capture was taken from (bench.py reports it next to roofline.traffic)
    summary["git"] = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except Exception:  # noqa: BLE001
    pass
with open(os.path.join(out_dir, f"ncu_{tag}_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)

lines = [f"# ncu summary ({tag})", "",
         "Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares):", "",
         "| kernel | launches | mean us | share |", "|---|---|---|---|"]
for k, v in summary["launches"].items():
    lines.append(f"| `{re.sub(r'[|]', '/', k)[:90]}` | {v['n']} | {v['mean_us']:.1f} | {v['share']:.3f} |")
lines += ["", "Full capture (`ncu --set full --clock-control none --import-source on`), per launch:", ""]
for name, ks in summary["kernels"].items():
    k = ks[-1]
    lines.append(f"## `{name}`  ({len(ks)} captured launch(es); last one shown)")
    lines.append("")
    for key in ("duration", "dram_read", "dram_write", "dram_bytes", "registers", "warp_instructions", "issue_active_pct",
                "warps_active_pct", "pipe_alu_pct", "pipe_fma_pct", "pipe_fp64_pct", "pipe_xu_pct", "pipe_lsu_pct",
                "pipe_tensor_pct", "pipe_tc_pct", "pipe_tc_cycles_pct", "pipe_tensor_hmma_pct", "pipe_tmem_pct",
                "stall_barrier", "stall_wait", "stall_math_pipe", "stall_no_instruction", "stall_long_scoreboard",
                "stall_short_scoreboard", "stall_not_selected", "stall_dispatch", "stall_branch_resolving", "pipe_fma_cycles_pct",
                "icache_hit_pct"):
        if key in k:
            lines.append(f"* {key}: {k[key]:.6g}")
    lines.append("")
with open(os.path.join(out_dir, f"ncu_{tag}_summary.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
