"""Summarise an `ncu --set full` capture for profiles/: for each kernel name keep the launch with the
largest grid (the 16-frame block launch), write its raw metrics as (kernel, metric, unit, value) rows and
the handful of numbers bench.py and DESIGN.md quote as JSON.

  ncu -i capture.ncu-rep --page raw --csv > raw.csv
  python tools/extract_traffic.py raw.csv profiles/ncu_block_<tag>_summary.csv profiles/traffic_<tag>.json [frames_per_launch]
"""
import csv
import json
import re
import sys

raw, out_csv, out_json = sys.argv[1:4]
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 16
rows = list(csv.reader(open(raw)))
header, units, body = rows[0], rows[1], rows[2:]
col = {name: i for i, name in enumerate(header)}


def num(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def short(name):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


best = {}
for r in body:
    if len(r) < len(header):
        continue
    name = short(r[col["Kernel Name"]])
    grid = 1
    for v in re.findall(r"\d+", r[col["Grid Size"]]):
        grid *= int(v)
    if name not in best or grid > best[name][0]:
        best[name] = (grid, r)

with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for name, (_, r) in sorted(best.items()):
        for h, u, v in zip(header, units, r):
            if num(v) is not None or h in ("Kernel Name", "Grid Size", "Block Size"):
                w.writerow([name, h, u, v])

KEYS = {
    "dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
    "gpu_time_us_under_ncu": "gpu__time_duration.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram_cycles_active_pct": "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum", "registers_per_thread": "launch__registers_per_thread",
    "pipe_fmaheavy_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "stall_math_pipe_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}
summary = {}
for name, (grid, r) in sorted(best.items()):
    entry = {"kernel": r[col["Kernel Name"]][:90], "grid": r[col["Grid Size"]], "frames_per_launch": frames}
    for key, metric in KEYS.items():
        if metric in col:
            value = num(r[col[metric]])
            if value is not None:
                value *= SCALE.get(units[col[metric]], 1.0)
            entry[key] = value
    if entry.get("dram_bytes_read") is not None and entry.get("dram_bytes_write") is not None:
        entry["dram_bytes_per_launch"] = entry["dram_bytes_read"] + entry["dram_bytes_write"]
    entry["source"] = out_csv + " (ncu --set full --clock-control none; the launch with the largest grid per kernel)"
    summary[name] = entry
json.dump(summary, open(out_json, "w"), indent=1)
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk not in ("kernel", "source")} for k, v in summary.items()}, indent=1))
