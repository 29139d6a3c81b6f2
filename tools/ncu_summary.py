#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel capture, --set full) into a few lines of markdown:
duration, registers, occupancy, pipe utilisation, DRAM traffic, top stall reasons and the
dynamic SASS opcode mix.  Usage: python tools/ncu_summary.py <file.ncu-rep> [units_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0][1] if rows and len(rows[0]) > 1 else "?", rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    m = raw(rep)
    name, hdr, data = source(rep)
    g = lambda k: m.get(k, ("", "n/a"))  # noqa: E731
    print(f"### `{name}`  ({rep.split('/')[-1]})\n")
    keys = [
        ("gpu__time_duration.sum", "duration"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers/thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ]
    print("| metric | value |\n|---|---|")
    for k, label in keys:
        u, v = g(k)
        print(f"| {label} (`{k}`) | {v} {u} |")
    stalls = []
    for k, (u, v) in m.items():
        mm = re.match(r"smsp__pcsamp_warps_issue_stalled_([a-z_]+)$", k)
        if mm and not k.endswith("_not_issued"):
            try:
                stalls.append((float(v.replace(",", "")), mm.group(1)))
            except ValueError:
                pass
    tot = sum(s for s, _ in stalls) or 1.0
    print("\nTop warp-state samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:6]))
    i_s, i_e = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    for r in data:
        mm = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[i_s])
        ops[mm.group(2) if mm else "?"] += int(r[i_e])
    total = sum(ops.values()) or 1
    print("\nDynamic SASS mix: " + ", ".join(f"{o} {100 * c / total:.1f}%" for o, c in ops.most_common(14)))
    if len(sys.argv) > 2:
        units = float(sys.argv[2])
        print(f"\nWarp instructions per unit: {total / units:.2f} (thread-level: {32 * total / units:.1f})")
    print()


if __name__ == "__main__":
    main()
