#!/usr/bin/env python
"""Aggregates the warp-stall samples and executed warp-instructions of an `ncu --set full --import-source on` capture of ba_point_kernel
per PHASE of the kernel (source-line ranges of dm-vio_b200/csrc/ba_point.cu).  Input: `ncu -i <rep> --page source --csv --print-source cuda,sass`.

    ncu -i gpurun_out/prof_ba_point.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_mix.csv
    python tools/phase_samples.py /tmp/src_mix.csv
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    rows = list(csv.reader(open(path)))
    cur, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; continue
        if r[0] and r[0].isdigit() and hdr:
            extra = len(r) - len(hdr)              # an unescaped quote inside the source text splits the field: fold the pieces back
            fields = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
            try:
                data.append((os.path.basename(cur), int(fields[0]), fields[1], int(fields[4] or 0), int(fields[7] or 0)))
            except ValueError:
                continue
    tot_s, tot_i = sum(d[3] for d in data), sum(d[4] for d in data)
    src = open(os.path.join(ROOT, "dm-vio_b200", "csrc", "ba_point.cu")).read().split("\n")

    def find(pat):
        for i, l in enumerate(src):
            if pat in l:
                return i + 1
        raise SystemExit(f"marker not found: {pat}")

    marks = [("prologue: chunk decode, adjoint cp.async, prior", find("ba_point_kernel(const __grid_constant__")),
             ("A: point / state loads (+ fused resubstitute)", find("// ---- direct loads")),
             ("A: centre + pattern projection", find("// ---- centre pixel at the FEJ point")),
             ("A: 4-tap float4 gather + interpolation", find("float h0 = 0.f, h1 = 0.f, h2 = 0.f;")),
             ("A: residual, Huber weight, 15 eight-lane sums", find("// ---- photometric residual, gradient weight, Huber")),
             ("A: classification + per-residual outputs", find("// ---- classification")),
             ("A: Jacobians, per-point record, 13x13 rows -> smem", find("      if (in) {")),
             ("A: warp counters", find("    float es = e_sum")),
             ("barrier + A': fold group partials, fp64 REDs", find("  cp_async_wait_all();  // the adjoint blocks")),
             ("B: point finalisation + Schur vectors", find("// ---------------------------------------------------------------- phase B")),
             ("C: Gram tiles + fp64 REDs", find("// ---------------------------------------------------------------- phase C")),
             ("end", find("static void launch_cfg"))]
    agg = collections.OrderedDict((m[0], [0, 0]) for m in marks[:-1])
    other = [0, 0]
    for f, l, s, sa, ins in data:
        k = next((i for i in range(len(marks) - 1) if f == "ba_point.cu" and marks[i][1] <= l < marks[i + 1][1]), None)
        tgt = agg[marks[k][0]] if k is not None else other
        tgt[0] += sa; tgt[1] += ins
    print(f"{tot_s} stall samples, {tot_i} executed warp-instructions (one launch, 129 CTAs x 28 warps)\n")
    print("| phase (source-line range of `ba_point.cu`) | stall samples | share | warp-instructions | share |")
    print("|---|---|---|---|---|")
    for k, (sa, ins) in agg.items():
        print(f"| {k} | {sa} | {100 * sa / tot_s:.1f} % | {ins} | {100 * ins / tot_i:.1f} % |")
    print(f"| inlined helpers of `ba_common.cuh` (8-lane shuffle sums, `pick8`, cp.async, RED wrappers: called from A, A', C) | {other[0]} | "
          f"{100 * other[0] / tot_s:.1f} % | {other[1]} | {100 * other[1] / tot_i:.1f} % |")
    print("\nTop source lines by stall samples:\n")
    print("| file:line | samples | warp-instr | source |")
    print("|---|---|---|---|")
    for f, l, s, sa, ins in sorted(data, key=lambda d: -d[3])[:12]:
        print(f"| {f}:{l} | {sa} | {ins} | `{s.strip()[:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
