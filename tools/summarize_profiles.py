#!/usr/bin/env python
"""Turns what a GPU visit left in gpurun_out/ into the committed evidence under profiles/ (named per round).

  usage: python tools/summarize_profiles.py r01 [tag]
  in : gpurun_out/launches.csv            ncu --metrics gpu__time_duration.sum --clock-control none  (launch list)
       gpurun_out/prof_<kernel>.ncu-rep   ncu --set full --clock-control none --import-source on     (one capture per kernel)
       gpurun_out/bench1.json, bench_ref.json, pytest_gpu.log, gpu.txt
  out: profiles/<round>[_tag]_launches.md / _launches.csv, profiles/<round>[_tag]_<kernel>_raw.csv + _summary.md, copies of the bench lines
"""
import collections
import csv
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def launches(prefix):
    src = os.path.join(G, "launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, rows = rows[hdr], rows[hdr + 1:]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(prefix + "_launches.md", "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:70]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |\n")
        step = {k: sum(v) / len(v) for k, v in d.items() if "l2_flush" not in k and "repack" not in k and "make_dI" not in k}
        st = sum(step.values())
        f.write("\nShare of one GN step (flush / upload kernels excluded): " + ", ".join(f"`{k.split('(')[0][:40]}` {100 * v / st:.1f}%" for k, v in step.items()) + "\n")
    shutil.copy(src, prefix + "_launches.csv")


def reports(prefix):
    for rep in sorted(glob.glob(os.path.join(G, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(f"{prefix}_{name}_raw.csv", "w").write(raw)
        rows = list(csv.reader(raw.splitlines()))
        h, units = rows[0], rows[1]
        with open(f"{prefix}_{name}_summary.md", "w") as f:
            f.write(f"`ncu --set full --clock-control none --import-source on` of `{name}` (file {os.path.basename(rep)}; full raw page in {os.path.basename(prefix)}_{name}_raw.csv)\n\n")
            for r in rows[2:]:
                f.write(f"### launch id {r[0]}: `{r[h.index('Kernel Name')][:80]}`\n\n| metric | value | unit |\n|---|---|---|\n")
                for k in KEY:
                    if k in h:
                        f.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
                f.write("\n")
        # where the time goes: warp-stall samples and executed instructions between consecutive block barriers of the SASS (phase boundaries)
        sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        rows = [r for r in csv.reader(sass.splitlines())]
        hi = next((i for i, r in enumerate(rows) if "# Samples" in r), None)
        if hi is not None:
            h2, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
            iS, iI, isrc = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Source")
            stall = [c for c in h2 if c.startswith("stall_") and "Not Issued" not in c]
            tot, toti = sum(int(r[iS]) for r in data) or 1, sum(int(r[iI]) for r in data) or 1
            with open(f"{prefix}_{name}_summary.md", "a") as f:
                f.write(f"### segments between block barriers ({tot} warp-stall samples, {toti} warp instructions; first kernel of the report)\n\n")
                f.write("| SASS index | ends at | samples | share | instructions share | top stall reasons |\n|---|---|---|---|---|---|\n")
                acc = acci = 0
                agg = collections.Counter()
                for k, r in enumerate(data):
                    acc += int(r[iS]); acci += int(r[iI])
                    for c in stall:
                        agg[c] += int(r[h2.index(c)] or 0)
                    src = r[isrc].strip()
                    if src.startswith("BAR.") or "ATOMG" in src or src.startswith("EXIT"):
                        top = ", ".join(f"{c[6:]} {100 * v / max(1, sum(agg.values())):.0f}%" for c, v in agg.most_common(3))
                        f.write(f"| {k} | `{src[:40]}` | {acc} | {100 * acc / tot:.1f}% | {100 * acci / toti:.1f}% | {top} |\n")
                        acc = acci = 0
                        agg = collections.Counter()
                    if src.startswith("EXIT") and k > len(data) // 2:
                        break
                f.write("\n")


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    tag = ("_" + sys.argv[2]) if len(sys.argv) > 2 else ""
    os.makedirs(P, exist_ok=True)
    prefix = os.path.join(P, rnd + tag)
    launches(prefix)
    reports(prefix)
    for f in sorted(glob.glob(os.path.join(G, "ev_*.json")) + glob.glob(os.path.join(G, "ev_pytest*.log")) + glob.glob(os.path.join(G, "scale_*.json"))):
        shutil.copy(f, prefix + "_" + os.path.basename(f)[3:] if os.path.basename(f).startswith("ev_") else prefix + "_" + os.path.basename(f))


if __name__ == "__main__":
    main()
