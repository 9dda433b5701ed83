#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep: python tools/ncu_summary.py FILE.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64inst%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64pipe%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dadd"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "dmul"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "dfma"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_ld_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local_st_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "glob_ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "glob_ld_reqs"),
]
STALL2 = "smsp__average_warps_issue_stalled_"


def main(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(p, "no data")
            continue
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(h, r))
            u = dict(zip(h, units))
            print(f"=== {p}: {d.get('Kernel Name', '')[:100]}")
            for k, lab in KEYS:
                if k in d:
                    print(f"  {lab:18s} {d[k]:>16s} {u[k]}")
            st = [(float(v.replace(',', '')), k) for k, v in d.items() if (k.startswith(STALL2) and k.endswith("_per_warp_active.pct") and "not_issued" not in k and v)]
            for v, k in sorted(st, reverse=True)[:7]:
                print(f"  stall {k[len(STALL2):-len('_per_warp_active.pct')]:28s} {v:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1:])
