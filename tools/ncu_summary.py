#!/usr/bin/env python
"""Print the key metrics of every kernel in an .ncu-rep (needs ncu on PATH): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp16.sum",
    "sm__cycles_elapsed.avg", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        print("==", path, r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:80s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
        st = [(float(r[i] or 0), h[len(STALL):]) for i, h in enumerate(hdr) if h.startswith(STALL) and not h.endswith("_not_issued")]
        tot = sum(v for v, _ in st) or 1.0
        print("  stalls:", ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:7]))
