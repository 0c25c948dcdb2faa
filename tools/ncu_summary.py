#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics (raw page) + instruction mix by source line (source page).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-70s %s %s" % (w, vals[i], units[i]))
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    # find header row
    h = None
    for i, r in enumerate(src):
        if "Instructions Executed" in r:
            h = i
            break
    if h is None:
        return
    hd = src[h]
    iex, ist = hd.index("Instructions Executed"), hd.index("Warp Stall Sampling (All Samples)")
    isrc = hd.index("Source")
    tot = sum(int(r[iex] or 0) for r in src[h + 1:] if len(r) > iex and (r[iex] or "0").isdigit())
    tst = sum(int(r[ist]) for r in src[h + 1:] if len(r) > ist and (r[ist] or "").isdigit())
    print("total warp instructions", tot, "stall samples", tst)
    # opcode mix
    mix = defaultdict(int)
    stall = defaultdict(int)
    for r in src[h + 1:]:
        if len(r) <= iex or not (r[iex] or "0").isdigit():
            continue
        toks = r[isrc].split()
        op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
        op = op.split(".")[0]
        mix[op] += int(r[iex] or 0)
        stall[op] += int(r[ist]) if (r[ist] or "").isdigit() else 0
    print("opcode mix (share of executed warp instructions | share of stall samples):")
    for op, c in sorted(mix.items(), key=lambda kv: -kv[1])[:nlines]:
        print("  %-10s %6.2f%%  | %6.2f%%" % (op, 100.0 * c / max(1, tot), 100.0 * stall[op] / max(1, tst)))


if __name__ == "__main__":
    main()
