#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` into {metric: [value per launch]} for the metrics profiles/ cites."""
import csv
import json
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]


def summarise(rep: str) -> dict:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, body = rows[hi], rows[hi + 1], rows[hi + 2:]
    res = {}
    for c, name in enumerate(hdr):
        if any(name == k or name.startswith(k + ".") and name in KEEP or name == k for k in KEEP):
            res[f"{name} [{units[c]}]"] = [r[c][:60] for r in body if len(r) == len(hdr)]
    return res


if __name__ == "__main__":
    print(json.dumps({a: summarise(a) for a in sys.argv[1:]}, indent=1))
