#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU needed) into the few numbers DESIGN.md and bench.py quote.

    python profiles/ncu_summary.py gpurun_out/prof_x.ncu-rep [...]  > profiles/rNN_<what>.txt
    python profiles/ncu_summary.py --stalls gpurun_out/prof_x.ncu-rep     # + top stall reasons per source line
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_warps", "occ_lim_warps"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct2"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1_global_ld_bytes"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_sb"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall_short_sb"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall_barrier"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall_lg_throttle"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "stall_mio_throttle"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall_math_throttle"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "stall_wait"),
    ("smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "stall_no_inst"),
    ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "stall_branch"),
    ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "stall_membar"),
    ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "stall_not_selected"),
    ("smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "stall_sleeping"),
    ("smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "stall_dispatch"),
    ("smsp__warp_issue_stalled_drain_per_warp_active.pct", "stall_drain"),
]


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarise(path):
    hdr, units, rows = raw_rows(path)
    print("== %s" % path)
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        print("-- %s" % name[:110])
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                print("   %-24s %s %s" % (label, r[i], units[i]))
        # warp-state sampling: share of each stall reason
        pre = "smsp__pcsamp_warps_issue_stalled_"
        samp = [(c[len(pre):], float(r[i])) for i, c in enumerate(hdr) if c.startswith(pre) and not c.endswith("not_issued")]
        tot = sum(v for _, v in samp) or 1.0
        print("   stall reasons (share of samples): " + ", ".join(
            "%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(samp, key=lambda kv: -kv[1]) if v / tot > 0.03))


def opcodes(path, top=14):
    """stall samples and executed warp instructions by SASS opcode (first kernel of the report)"""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, tab = None, []
    for r in rows:
        if "Source" in r and hdr is None:
            hdr = r
            continue
        if hdr and "Source" in r:
            break  # next kernel
        if hdr and len(r) == len(hdr):
            tab.append(r)
    if not hdr:
        return
    si, ci, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")

    def num(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    st, ins = {}, {}
    for r in tab:
        parts = r[si].split()
        if not parts:
            continue
        op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
        st[op] = st.get(op, 0.0) + num(r[ci])
        ins[op] = ins.get(op, 0.0) + num(r[ii])
    ts, ti = sum(st.values()) or 1.0, sum(ins.values()) or 1.0
    print("   by opcode (stall-sample share / executed-instruction share), %d warp instructions:" % ti)
    for op, v in sorted(st.items(), key=lambda kv: -kv[1])[:top]:
        print("     %-8s %5.1f%% / %5.1f%%" % (op, 100.0 * v / ts, 100.0 * ins[op] / ti))


def stalls(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the source page prints one table per kernel; keep rows that have a sampling column
    hdr = None
    table = []
    for r in rows:
        if "Source" in r and any("Samples" in c for c in r):
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            table.append(r)
    if not hdr:
        print("no source page (compile with -lineinfo, capture with --import-source on)")
        return
    si = hdr.index("Source")
    ci = [i for i, c in enumerate(hdr) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)")]
    ci = ci[0] if ci else [i for i, c in enumerate(hdr) if "Samples" in c][0]

    def num(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    table.sort(key=lambda r: -num(r[ci]))
    tot = sum(num(r[ci]) for r in table) or 1.0
    print("   top stall-sample lines (%s):" % hdr[ci])
    for r in table[:top]:
        print("   %6.2f%%  %s" % (100.0 * num(r[ci]) / tot, r[si].strip()[:150]))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    for p in args:
        summarise(p)
        if "--ops" in sys.argv:
            opcodes(p)
        if "--stalls" in sys.argv:
            stalls(p)
