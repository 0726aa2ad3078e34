#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) into the handful of metrics DESIGN.md / bench.py cite.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rN_x_ncu.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s" % r[head.index("Kernel Name")])
        for name in WANT:
            if name in head:
                i = head.index(name)
                print("  %-70s %s %s" % (name, r[i], units[i]))
        stalls = []
        for i, name in enumerate(head):
            if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("  stall reasons (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:8]))
        if "dram__bytes_read.sum" in head:
            try:
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(r[head.index("dram__bytes_read.sum")]) * mult.get(units[head.index("dram__bytes_read.sum")], 1)
                wr = float(r[head.index("dram__bytes_write.sum")]) * mult.get(units[head.index("dram__bytes_write.sum")], 1)
                print("  dram traffic per launch (bytes): %.0f" % (rd + wr))
            except ValueError:
                pass


if __name__ == "__main__":
    main(sys.argv[1])
