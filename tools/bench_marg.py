#!/usr/bin/env python
"""Secondary benchmark: the point marginalisation of makeKeyFrame — FullSystem::flagPointsForRemoval's
linearize / applyRes / fixLinearizationF loop + EnergyFunctional::marginalizePointsF for a third of a 7 KF / 2000 point window
(both single-threaded in the reference).  GPU: one dmv_ba_marginalize_points call (mask + table upload, MARG point kernel, stitch,
D2H of the system and the residual states, host unpack) through ctypes; CPU: the oracle's fixLinearization + marginalizePoints (fp32
three-tier accumulators like the reference), rebuilt per repetition because the call mutates the window.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--frac", type=float, default=0.33)
    args = ap.parse_args()
    import dmvio_b200.capi as capi
    import dmvio_b200.synth as synth
    from oracle import orc
    from helpers import product_ba_from_oracle, rel
    W = synth.make_window(nf=7, npts=2000, seed=1234)
    rng = np.random.default_rng(1)
    npts = len(W["host"])
    pts = np.sort(rng.choice(npts, int(npts * args.frac), replace=False)).astype(np.int32)
    ow = orc.Window(W)
    ba = product_ba_from_oracle(capi, W, ow)
    ba.linearize(); ba.apply_res()
    ad, cd = ow.adHTdeltaF(), ow.calib()["cDeltaF"]
    for _ in range(5):
        g = ba.marginalize_points(pts, ad, cd)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        g = ba.marginalize_points(pts, ad, cd)
    gpu_ms = (time.perf_counter() - t0) / args.reps * 1e3
    cpu_reps = max(3, args.reps // 20)
    wins = [orc.Window(W) for _ in range(cpu_reps)]
    t0 = time.perf_counter()
    for wv in wins:
        o = wv.marginalize(pts, precision=0)
    cpu_ms = (time.perf_counter() - t0) / cpu_reps * 1e3
    d = orc.Window(W).marginalize(pts, precision=1)
    print(json.dumps({"metric": "point marginalisation ms/keyframe (7 KF, %d of %d points, %d residuals, 640x480)" % (len(pts), npts, int(d["resInM"])),
                      "gpu_ms": gpu_ms, "cpu_oracle_ms": cpu_ms, "cpu_threads": 1, "speedup": cpu_ms / gpu_ms,
                      "rel_err_H_vs_fp64_oracle": rel(g["H"], d["H"]), "rel_err_b_vs_fp64_oracle": rel(g["b"], d["b"]),
                      "resInM_gpu": int(g["resInM"]), "resInM_cpu": int(d["resInM"]),
                      "timed_gpu": "dmv_ba_marginalize_points through ctypes: uploads, 2 memsets, MARG point kernel + stitch, D2H of system / states / res_toZeroF, unpack"}))
    ba.close()


if __name__ == "__main__":
    main()
