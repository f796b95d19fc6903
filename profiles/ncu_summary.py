#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page) into the few numbers the
roofline discussion needs.  Usage: ncu_summary.py prof.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki][:70])
        for w in WANT:
            if w in hdr:
                print("  %-86s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # first kernel only
    hdr = rows[1]
    ci, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == "Address":
            if data:
                break
            continue
        try:
            data.append((r[ci], int(r[ie] or 0), int(r[ws] or 0), [int(r[i] or 0) for i in stall_cols]))
        except ValueError:
            pass
    ti, tw = sum(d[1] for d in data), sum(d[2] for d in data)
    print("== source page, first launch: %d SASS lines, %d warp-instructions, %d stall samples" % (len(data), ti, tw))
    tot = [0] * len(stall_cols)
    for d in data:
        for k, v in enumerate(d[3]):
            tot[k] += v
    order = sorted(range(len(stall_cols)), key=lambda k: -tot[k])
    print("  stall reasons:", ", ".join("%s %.1f%%" % (hdr[stall_cols[k]][6:], 100.0 * tot[k] / max(1, sum(tot))) for k in order[:8]))
    print("  hottest instructions by stall samples:")
    for k, d in sorted(enumerate(data), key=lambda x: -x[1][2])[:14]:
        print("   %4d %-64s inst %5.2f%% stall %5.2f%%" % (k, d[0][:64], 100.0 * d[1] / ti, 100.0 * d[2] / max(1, tw)))


if __name__ == "__main__":
    main()
