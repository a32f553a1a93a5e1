#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) as the markdown table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/long_fwd_r02.ncu-rep [launch index] > table.md
Reads `ncu -i <rep> --page raw --csv`; no GPU needed."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, data = rows[0], rows[1], rows[2:]
    row = data[which]
    col = {h: i for i, h in enumerate(head)}
    print(f"kernel: `{row[col['Kernel Name']]}`  grid {row[col['Grid Size']]} block {row[col['Block Size']]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEEP:
        if k in col:
            print(f"| {k} | {row[col[k]]} | {units[col[k]]} |")
    stalls = [(h, row[i]) for h, i in col.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    def num(v):
        try:
            return float(v.replace(",", ""))
        except ValueError:
            return 0.0
    stalls.sort(key=lambda kv: -num(kv[1]))
    for h, v in stalls[:9]:
        print(f"| {h} | {v} | samples |")


if __name__ == "__main__":
    main()
