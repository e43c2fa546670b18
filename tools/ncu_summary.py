#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md argues from.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
    "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_tma_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
]
STALLS = "smsp__average_warp_latency_issue_stalled_"
STALLS2 = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} kernel launch(es) captured with ncu --set full --clock-control none")
    for r in rows[2:]:
        print()
        print("kernel:", r[col["Kernel Name"]])
        for k in KEYS:
            if k in col and r[col[k]] != "":
                print(f"  {k:70s} {r[col[k]]} {units[col[k]]}")
        st = []
        for k, i in col.items():
            if (k.startswith(STALLS) or k.startswith(STALLS2)) and k.endswith("_per_warp_active.pct") is False and k.endswith(".ratio"):
                try:
                    st.append((float(r[i]), k))
                except ValueError:
                    pass
        st.sort(reverse=True)
        if st:
            print("  top stall reasons (avg warps stalled per issue cycle / latency ratio):")
            for v, k in st[:8]:
                print(f"    {k:74s} {v:.3f}")


if __name__ == "__main__":
    main()
