#!/usr/bin/env python3
"""Print the handful of ncu metrics DESIGN.md quotes, per kernel, from a --set full report.

    python tools/ncu_summary.py REPORT.ncu-rep [title]"""
import csv
import subprocess
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__icc_request_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("ncu --set full --clock-control none, %s  %s" % (d.get("Kernel Name", "?").split("(")[0], title))
        for k in WANT:
            if k in d:
                print("  %s [%s] = %s" % (k, u[k], d[k]))
        st = {k[len(STALL):]: float(v.replace(",", "") or 0) for k, v in d.items() if k.startswith(STALL) and not k.endswith("_not_issued") and v}
        tot = sum(st.values())
        if tot > 0:
            print("  stall reasons (pc sampling, share of samples):")
            for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
                print("    %6.2f%% %s" % (100.0 * v / tot, k))
        print()


if __name__ == "__main__":
    main()
