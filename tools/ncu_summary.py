#!/usr/bin/env python
"""Condenses `ncu -i <rep> --page raw --csv` into the per-kernel summary kept under profiles/ (one block per distinct
kernel: the launch with the longest duration).  usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(sys.stdin))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    best = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = r[col["Kernel Name"]].split("(")[0]
        t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        if k not in best or t > best[k][0]:
            best[k] = (t, r)
    for k, (t, r) in best.items():
        print(f"===  {k} grid {r[col['Grid Size']]}")
        for m in KEEP:
            if m in col:
                print(f"   {m:<78} {r[col[m]]:>16} {units[col[m]]}")
        st = []
        for n, i in col.items():
            if n.startswith(STALL) and n.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i].replace(",", "")), n[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        st.sort(reverse=True)
        print("   stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:7]))


if __name__ == "__main__":
    main()
