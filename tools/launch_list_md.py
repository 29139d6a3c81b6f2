#!/usr/bin/env python
"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel share table.
Usage: python tools/launch_list_md.py <launches.csv> [title]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], ms))
    tot = sum(ms for _, ms in rows) or 1.0
    agg = collections.OrderedDict()
    for name, ms in rows:
        key = re.sub(r"\(.*$", "", name)[:80]
        a = agg.setdefault(key, [0, 0.0, []])
        a[0] += 1; a[1] += ms; a[2].append(ms)
    print(f"# ncu launch list, {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (first 400 launches of the run; per-launch "
          "times are cold-cache and serialised: compare shares, not absolutes).\n")
    print("| kernel | launches | total ms | avg ms | max ms | share |\n|---|---|---|---|---|---|")
    for key, (cnt, ms, lst) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"| `{key}` | {cnt} | {ms:.3f} | {ms / cnt:.4f} | {max(lst):.4f} | {100 * ms / tot:.1f}% |")
    ours = sum(ms for name, ms in rows if "pnp::" in name)
    print(f"\nOur kernels (`pnp::*`) account for {100 * ours / tot:.1f}% of the device time in the capture; the rest is "
          "torch generating the synthetic inputs (outside every timed step).")


if __name__ == "__main__":
    main()
