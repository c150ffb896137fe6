#!/usr/bin/env python
"""Key metrics of the launches in an .ncu-rep (read on the CPU box):  python tools/ncu_keys.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    print(f"  {k} = {r[i][:110]} {units[i]}")
        print()


if __name__ == "__main__":
    main()
