#!/usr/bin/env python
"""Summarise an `ncu --set full` report for profiles/:
    python tools/ncu_summary.py gpurun_out/r2_prof.ncu-rep r2
writes profiles/<tag>_ncu_full_summary.json (selected raw metrics per kernel, last launch of each
name), profiles/<tag>_ncu_<kernel>_details.txt (the details page) and profiles/ncu_traffic.json
(dram bytes per launch + the hash of each kernel's source file: bench.py quotes the traffic only
while the source still matches what was profiled)."""
import csv
import hashlib
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rep, tag = sys.argv[1], sys.argv[2]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
SOURCES = {"sync_pair_kernel": "ragnar_b200/csrc/rgc_sync_pair.cu",
           "sync_prologue_kernel": "ragnar_b200/csrc/rgc_sync_pair.cu",
           "sync_sort_kernel": "ragnar_b200/csrc/rgc_sync_pair.cu",
           "energy_hist_kernel": "ragnar_b200/csrc/rgc_hist_device.cuh",
           "sync_literal_kernel": "ragnar_b200/csrc/rgc_sync_literal.cu"}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
kcol = hdr.index("Kernel Name")
summary = {}
for r in data:
    if len(r) != len(hdr):
        continue
    ent = {}
    for m in KEEP:
        if m in hdr:
            ent[m] = {"value": r[hdr.index(m)], "unit": units[hdr.index(m)]}
    summary[r[kcol]] = ent  # the last launch of a name wins
(ROOT / "profiles" / f"{tag}_ncu_full_summary.json").write_text(json.dumps(summary, indent=1))


def to_bytes(ent):
    v, u = float(ent["value"].replace(",", "")), ent["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


traffic = {"capture": f"profiles/{tag}_ncu_full_summary.json", "kernels": {}}
for name, ent in summary.items():
    short = re.sub(r"^void ", "", name).split("<")[0].split("(")[0]
    if short in SOURCES and "dram__bytes_read.sum" in ent:
        src = ROOT / SOURCES[short]
        traffic["kernels"][short] = {
            "dram_bytes": to_bytes(ent["dram__bytes_read.sum"]) + to_bytes(ent["dram__bytes_write.sum"]),
            "source": SOURCES[short], "source_sha16": hashlib.sha256(src.read_bytes()).hexdigest()[:16],
            "particles": int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000,
            "bins": int(sys.argv[4]) if len(sys.argv) > 4 else 200}
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--kernel-name", f"regex:^{re.escape(short)}"],
                         capture_output=True, text=True).stdout
    if det.strip() and short in SOURCES:
        # keep the last launch's sections only
        blocks = det.split("\n  " + name.split("(")[0])
        text = det if len(blocks) < 2 else "  " + name.split("(")[0] + blocks[-1]
        (ROOT / "profiles" / f"{tag}_ncu_{short}_details.txt").write_text(text)
(ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
print(json.dumps({k: {m: v["value"] + " " + v["unit"] for m, v in e.items() if m in KEEP[:4]}
                  for k, e in summary.items()}, indent=1))
