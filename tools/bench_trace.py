#!/usr/bin/env python
"""Secondary benchmark: FullSystem::traceNewCoarse — ImmaturePoint::traceOn for the immature points of every host
keyframe against the newest frame (runs on EVERY tracked frame in the reference, single-threaded under mapMutex).
GPU: one dmv_ct_trace_points call per host frame (H2D of the point arrays, kernel, D2H of the 7 in/out words) on the frame already
resident in the coarse-tracker handle; CPU: the oracle's traceOn loop (bit-identical results, asserted).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hosts", type=int, default=6)
    ap.add_argument("--points-per-host", type=int, default=1500)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    import dmvio_b200.capi as capi
    import dmvio_b200.hostmath as hm
    import dmvio_b200.synth as synth
    from oracle import orc
    nf = args.hosts + 1
    W = synth.make_window(nf=nf, npts=10, seed=5, trans=0.05, rot=0.01)
    w, h = W["w"], W["h"]
    rng = np.random.default_rng(3)
    new = nf - 1
    sets = []
    for host in range(args.hosts):
        u, v = rng.integers(10, w - 10, args.points_per_host), rng.integers(10, h - 10, args.points_per_host)
        sets.append((orc.ip_init(W["dI"][host], w, h, u, v), hm.trace_tables(W, host, new)))
    g = capi.CT(w, h, synth.pyr_levels(w, h), max_points=1024)
    g.upload_new(0, W["dI"][new])
    for P, (KRKi, Kt, aff) in sets:
        g.trace_points(P, KRKi, Kt, aff)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        outs1 = [g.trace_points(P, KRKi, Kt, aff) for P, (KRKi, Kt, aff) in sets]
    gpu_ms_per_host_calls = (time.perf_counter() - t0) / args.reps * 1e3
    msets = [(P, KRKi, Kt, aff) for P, (KRKi, Kt, aff) in sets]
    g.trace_points_multi(msets)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        outs = g.trace_points_multi(msets)
    gpu_ms = (time.perf_counter() - t0) / args.reps * 1e3
    t0 = time.perf_counter()
    cpu_reps = max(1, args.reps // 10)
    for _ in range(cpu_reps):
        refs = [orc.ip_trace(P, W["dI"][new], w, h, KRKi, Kt, aff) for P, (KRKi, Kt, aff) in sets]
    cpu_ms = (time.perf_counter() - t0) / cpu_reps * 1e3
    exact = all(np.array_equal(a[k], b[k], equal_nan=True) for a, b in zip(outs, refs) for k in orc.IP_STATE_KEYS)
    npts = args.hosts * args.points_per_host
    hist = np.bincount(np.concatenate([r["status"] for r in refs]), minlength=6).tolist()
    print(json.dumps({"metric": "traceNewCoarse ms/frame (%d hosts x %d immature points, %dx%d)" % (args.hosts, args.points_per_host, w, h),
                      "gpu_ms_per_frame": gpu_ms, "gpu_points_per_s": npts / (gpu_ms * 1e-3), "cpu_oracle_ms_per_frame": cpu_ms, "cpu_threads": 1,
                      "speedup": cpu_ms / gpu_ms, "gpu_ms_per_frame_one_call_per_host": gpu_ms_per_host_calls, "bit_identical_to_cpu": bool(exact), "status_histogram_GOOD_OOB_OUTLIER_SKIPPED_BADCOND_UNINIT": hist,
                      "timed_gpu": "all host frames in ONE dmv_ct_trace_points_multi call: H2D of 31 words/point, ip_trace_kernel (warp per point), D2H of 7 words/point, sync (through ctypes)"}))
    g.close()


if __name__ == "__main__":
    main()
