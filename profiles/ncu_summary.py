#!/usr/bin/env python
"""Prints the metrics of an .ncu-rep that the round summaries quote: python profiles/ncu_summary.py <file.ncu-rep> [regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = r"""gpu__time_duration.sum$|launch__registers_per_thread|launch__occupancy_limit|launch__shared_mem_per_block_dynamic|launch__grid_size
|sm__warps_active.avg.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active.avg.pct|smsp__issue_active.avg.pct_of_peak_sustained_active
|smsp__inst_executed.sum$|smsp__thread_inst_executed_per_inst_executed.ratio|dram__bytes_read.sum$|dram__bytes_write.sum$|lts__t_sector_hit_rate.pct
|l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$
|memory_l1_wavefronts_shared_ideal|l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg$|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio
|sm__inst_executed_pipe_(alu|fma|xu|lsu|fp64).avg.pct_of_peak_sustained_active|l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate""".replace("\n", "")


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else KEYS)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
        for h, u, v in zip(hdr, units, r):
            if pat.search(h):
                try:
                    fv = float(v.replace(",", ""))
                    if fv < 0.02 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                print(f"{h},{u},{v}")


if __name__ == "__main__":
    main()
