#!/usr/bin/env python
"""Summarise an .ncu-rep (read without a GPU): python tools/ncu_summary.py REPORT [ALGORITHMIC_BYTES]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    report = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    header = rows[0]
    units = rows[1]
    for data in rows[2:]:
        name = data[header.index("Kernel Name")]
        print("kernel {}  grid {} block {}".format(name, data[header.index("Grid Size")], data[header.index("Block Size")]))
        values = {}
        for column, unit, value in zip(header, units, data):
            if column in WANT:
                values[column] = (value, unit)
                print("  {:<82} {:>16} {}".format(column, value, unit))

        def number(key):
            value, unit = values[key]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
            return float(value.replace(",", "")) * scale.get(unit, 1.0)
        try:
            traffic = number("dram__bytes_read.sum") + number("dram__bytes_write.sum")
            seconds = number("gpu__time_duration.sum")
            print("  DRAM traffic per launch: {:.1f} MB  ({:.0f} GB/s over {:.1f} us)".format(
                traffic / 1e6, traffic / seconds / 1e9, seconds * 1e6))
            if len(sys.argv) > 2:
                algorithmic = float(sys.argv[2])
                print("  algorithmic bytes {:.1f} MB -> traffic/algorithmic = {:.3f}".format(
                    algorithmic / 1e6, traffic / algorithmic))
        except Exception as error:
            print("  (no traffic summary: {})".format(error))


if __name__ == "__main__":
    main()
