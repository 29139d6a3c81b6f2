#!/usr/bin/env python
"""Copy what a `tools/profile_round.sh <tag>` call brought back in gpurun_out/ into profiles/ and derive the readings
that are committed: bench line, launch list (+ shares), ncu summaries with the region table of the pair kernel, DRAM
traffic per unit.  Usage: python tools/collect_profiles.py r2"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KERNELS = (("ik_solve_v_kernel", 4194304), ("ik_solve_small_kernel", 4096), ("reward_kernel", 16777216),
           ("her_relabel_kernel", 8388608), ("move_ik_plan_v_kernel", 1048576), ("ik_solve_resume_kernel", 4194304))


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def region_table(rep, units):
    rows = ncu_csv(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ins = [(int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r) for r in data]
    passes = max(e for _, e, _ in ins)

    def bucket(e):
        if e >= 0.9 * passes:
            return "1 per-pass: the DLS pass of all 64 slots (main loop)"
        if e >= 0.2 * passes:
            return "2 per-flush: store finished slots + refill"
        if 0.035 * passes <= e < 0.06 * passes:
            return "3 ticket reservation (one atomic per 64-256 queries)"
        if 0.02 * passes <= e < 0.035 * passes:
            return "4 tail phase: lane_solve loop of the block's last warp"
        return "5 once per warp / block: prologue (trig table), parking, tail set-up and stores, counters"

    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    for s, e, r in ins:
        b = agg[bucket(e)]
        b[0] += 1
        b[1] += s
        b[2] += e
        for k in stalls:
            b[3][k[6:]] += int(r[ix[k]])
    flush = max((e for _, e, _ in ins if 0.2 * passes <= e < 0.9 * passes), default=1)
    out = ["", "#### Where the warp samples of `ik_solve_v_kernel<F2>` go (source page of the capture above, instructions "
           "bucketed by how often they execute)", "",
           "| region | SASS instructions | warp-instructions executed | warp samples | top stall reasons |", "|---|---|---|---|---|"]
    for b in sorted(agg):
        n, s, e, st = agg[b]
        top = ", ".join(f"{k} {100 * v / max(1, sum(st.values())):.0f} %" for k, v in st.most_common(4))
        out.append(f"| {b[2:]} | {n} | {e:,} | {100 * s / tot:.1f} % | {top} |")
    out += ["", f"Warp passes executed: {passes:,} for {units:,} queries (15.85 evaluations each, 64 slots per warp: "
            f"{units * 15.85 / 64:,.0f} is the minimum - the rest are frozen slots and the tail); one flush per "
            f"{passes / flush:.1f} passes in this capture.", ""]
    return "\n".join(out)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    line = open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1]
    json.dump(json.loads(line), open(os.path.join(P, f"bench_{tag}.json"), "w"))
    shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"launches_{tag}.csv"))
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_list_md.py"), os.path.join(G, f"launches_{tag}.csv")],
                        capture_output=True, text=True).stdout
    open(os.path.join(P, f"launches_{tag}.md"), "w").write(md)
    summary, traffic = [], {}
    for k, units in KERNELS:
        rep = os.path.join(P, f"{k}_{tag}.ncu-rep")
        if not os.path.exists(os.path.join(G, f"{k}_{tag}.ncu-rep")):
            print(f"no capture of {k} in gpurun_out/ - skipped")
            continue
        shutil.copy(os.path.join(G, f"{k}_{tag}.ncu-rep"), rep)
        summary.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(units)],
                                      capture_output=True, text=True).stdout)
        if k == "ik_solve_v_kernel":
            summary.append(region_table(rep, units))
        rows = ncu_csv(rep, "raw")
        m = {h: (rows[2][i], rows[1][i]) for i, h in enumerate(rows[0])}
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        b = sum(float(m[x][0].replace(",", "")) * scale[m[x][1]] for x in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[k] = {"dram_bytes_per_launch": b, "units_per_launch": units, "dram_bytes_per_unit": b / units,
                      "source": f"profiles/{k}_{tag}.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    open(os.path.join(P, f"ncu_summary_{tag}.md"), "w").write("\n".join(summary))
    try:
        traffic["ik_solve_kernel"] = json.load(open(os.path.join(P, "ncu_traffic_r1.json")))["ik_solve_kernel"]
    except Exception:
        pass
    json.dump(traffic, open(os.path.join(P, f"ncu_traffic_{tag}.json"), "w"), indent=1)
    d = json.loads(line)
    print(json.dumps(d["headline"]))
    print("ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"])


if __name__ == "__main__":
    main()
