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
    # the bench step is ONE launch of the IK kernel: the full-size launches are the slowest ones of that kernel
    ik = [ms for name, ms in rows if "ik_solve_v_kernel<pnp_spec::F2" in name or "ik_solve_kernel<float" in name]
    if ik:
        full = [ms for ms in ik if ms > 0.5 * max(ik)]
        print(f"\nFull-size IK launches in the capture: {len(full)}, avg {sum(full) / len(full):.3f} ms under ncu (compare "
              "`roofline.kernel_ms` of the bench line, CUDA events, plain run): the timed step is one launch of this "
              "kernel, so its share of a step is 100 % in both.")
    ours = sum(ms for name, ms in rows if "pnp::" in name)
    print(f"\nOur kernels (`pnp::*`) account for {100 * ours / tot:.1f}% of the device time in the capture; the rest is "
          "torch generating the synthetic inputs (outside every timed step).")


if __name__ == "__main__":
    main()
