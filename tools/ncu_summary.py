#!/usr/bin/env python
"""Text summary of the metrics we cite from an `ncu --set full` report (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/full_r01_k_tile_stream.ncu-rep > profiles/r01_k_tile_stream.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_shared_loads",
        "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "derived__memory_l1_wavefronts_shared_excessive",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("# source:", rep)
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for w in WANT + stalls:
            if w in hdr:
                v = r[hdr.index(w)]
                try:
                    if w in stalls and float(v) < 1.0:
                        continue
                except ValueError:
                    pass
                print(f"  {w:88s} {v:>18s} {units[hdr.index(w)]}")


if __name__ == "__main__":
    main(sys.argv[1])
