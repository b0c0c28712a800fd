#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few metrics the roofline discussion needs:
    python profiles/tools/ncu_raw_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg", "sm__pipe_tensor_subpipe_hmma", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct", "smsp__pcsamp_warps_issue_stalled", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg"]


def main():
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for vals in rows[2:]:
        print("==", vals[ik][:100], "grid", vals[hdr.index("Grid Size")], "block", vals[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, vals):
            if any(k in h for k in KEYS) and "_not_issued" not in h and ".per_second" not in h and ".pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ""):
                print(f"  {h:85s} {u:12s} {v}")


if __name__ == "__main__":
    main()
