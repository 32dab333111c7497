#!/usr/bin/env python
"""profiles/ helper: condense an .ncu-rep (ncu --set full) into a small tab-separated text summary, one column per
captured launch:  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "header comment" > profiles/rN/ncu_x.txt"""
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
]


def main():
    rep, comment = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if comment:
        print("# " + comment)
    names = KEEP + sorted(h for h in hdr if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"))
    for name in names:
        if name in hdr:
            i = hdr.index(name)
            print("\t".join([name, units[i]] + [r[i] for r in data]))


if __name__ == "__main__":
    main()
