#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i … --page raw --csv`) or a launch-list CSV into a short text table.
Usage: tools/ncu_summary.py raw <file.ncu-rep> | launches <launches.csv>"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
           "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
           "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("==", r[ki].split("(")[0])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("   %-75s %s %s" % (m, r[i], units[i]))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[ui], v)
        name = r[ki].split("(")[0].split("::")[-1]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-28s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-28s %6d %12.1f %10.1f %7.3f" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))
    # the kernels of ONE bench step (the other launches belong to bench.py's aligner_stress / sequence sections and to
    # the e2e repitch): their shares are what roofline.kernel_share_of_step must agree with
    step = ("fast_nms_kernel", "compact_kernel", "blur_kernel", "describe_tile_kernel", "match_kernel",
            "select_strips_kernel", "select_strips_kernel<16>", "linearize_pairs_kernel")
    stot = sum(agg[k][1] for k in step if k in agg)
    if stot:
        print()
        print("kernels of the bench step only:")
        for k in step:
            if k in agg:
                print("%-28s %6d %12.1f %10.1f %7.3f" % (k, agg[k][0], agg[k][1], agg[k][1] / agg[k][0], agg[k][1] / stot))


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2])
