#!/usr/bin/env python3
"""Print a compact markdown summary of one kernel in an ncu report (the numbers the roofline fields are built from).
    python tools/ncu_summary.py <report.ncu-rep> [frames_in_launch]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: vals[hdr.index(k)] if k in hdr else "n/a"
unit = lambda k: units[hdr.index(k)] if k in hdr else ""
keys = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
print(f"kernel: `{get('Kernel Name')}`  (report `{rep}`)\n")
print("| metric | value |\n|---|---|")
for k in keys:
    if k in hdr:
        print(f"| {k} | {get(k)} {unit(k)} |")
if frames:
    rd = float(get("dram__bytes_read.sum")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit("dram__bytes_read.sum")]
    wr = float(get("dram__bytes_write.sum")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit("dram__bytes_write.sum")]
    t = float(get("gpu__time_duration.sum")) * {"ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}[unit("gpu__time_duration.sum")]
    print(f"\nper frame ({frames} frames in this launch): dram read {rd / frames:,.0f} B, write {wr / frames:,.0f} B, "
          f"traffic {(rd + wr) / frames:,.0f} B; {frames / t / 1e6:.3f} Mframes/s under ncu; "
          f"{float(get('smsp__inst_executed.sum')) / frames:,.0f} warp-instructions/frame")
    print(f"TRAFFIC_JSON {{\"dram_bytes_per_frame\": {(rd + wr) / frames:.1f}, \"read\": {rd / frames:.1f}, \"write\": {wr / frames:.1f}}}")
