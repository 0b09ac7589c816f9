#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a markdown table per kernel launch and a traffic json.
usage: ncu_summary.py report.ncu-rep [title] > profiles/xxx.md ; also prints a JSON line to stderr."""
import csv, json, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    seen, traffic = set(), {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        print(f"## {name}\n\n| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in col:
                print(f"| `{m}` | {r[col[m]]} | {units[col[m]]} |")
        print()

        def num(m):
            v, u = float(r[col[m]].replace(",", "")), units[col[m]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        traffic[name] = {"dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
                         "duration": r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]}
    sys.stderr.write(json.dumps(traffic, indent=1) + "\n")


if __name__ == "__main__":
    main()
