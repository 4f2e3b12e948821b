#!/usr/bin/env python
"""Secondary benchmark (BASELINE config 2): CoarseTracker::trackNewestCoarse on a 640x480 synthetic pair,
GPU (C++ host adapter over the C ABI: H2D of the new image, pyramid on the device, one fused calcRes+calcGSSSE launch per LM
evaluation, 8x8 solve on the host) vs the CPU oracle (single thread, like the reference).  Prints one JSON line.
  python tools/bench_coarse.py [--levels 0|5] [--frames 200]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=0, help="0 = what setGlobalCalib yields (4 at 640x480); 5 = forced, as BASELINE config 2 words it")
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--cpu-frames", type=int, default=20)
    args = ap.parse_args()
    import dmvio_b200.hostapi as hostapi
    import dmvio_b200.synth as synth
    from oracle import orc
    T = synth.make_tracking_pair(seed=4321, levels=args.levels)
    L = T["levels"]
    g = hostapi.CoarseTracker(T["w"], T["h"], T["K"], L)
    counts = g.set_ref(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    # setCoarseTrackingRef (makeCoarseDepthL0, once per keyframe): host build + upload of the lists vs everything on the device
    t_start = time.perf_counter()
    for _ in range(20):
        g.set_ref(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    set_ref_host_ms = (time.perf_counter() - t_start) / 20 * 1e3
    for _ in range(3):
        g.set_ref_device(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["img_ref"])
    t_start = time.perf_counter()
    for _ in range(50):
        counts_dev = g.set_ref_device(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["img_ref"])
    set_ref_dev_ms = (time.perf_counter() - t_start) / 50 * 1e3
    assert counts_dev == counts
    R0, t0 = np.eye(3), np.zeros(3)
    # the two halves of a frame separately: upload + device pyramid, then tracking only (the CPU side is timed the same way)
    for _ in range(3):
        g.set_new_image(T["img_new"]); g.track(R0, t0, 0.0, 0.0)
    t_start = time.perf_counter()
    for _ in range(args.frames):
        g.set_new_image(T["img_new"])
        g.track(R0, t0, 0.0, 0.0, coarsest=0, minRes=np.zeros(5))   # aborts after level 0's first evaluation: ~ the cost of upload + pyramid + 1 launch
    upload_ms = (time.perf_counter() - t_start) / args.frames * 1e3
    track_only = {}
    for mode in (False, True):
        g.set_new_image(T["img_new"])
        for _ in range(3):
            g.track(R0, t0, 0.0, 0.0, device_lm=mode)
        t_start = time.perf_counter()
        for _ in range(args.frames):
            g.track(R0, t0, 0.0, 0.0, device_lm=mode)
        track_only[mode] = (time.perf_counter() - t_start) / args.frames * 1e3
    t_start = time.perf_counter()
    for _ in range(5):
        orc.make_images(T["img_new"], T["K"], args.levels)
    cpu_pyr_ms = (time.perf_counter() - t_start) / 5 * 1e3
    timings = {}
    for mode in (False, True):
        for _ in range(5):
            g.set_new_image(T["img_new"]); r = g.track(R0, t0, 0.0, 0.0, device_lm=mode)
        t_start = time.perf_counter()
        ev = 0
        for _ in range(args.frames):   # per frame: upload + device pyramid + the whole LM loop, like FullSystem::trackNewCoarse does per camera frame
            g.set_new_image(T["img_new"])
            r = g.track(R0, t0, 0.0, 0.0, device_lm=mode)
            ev += r["evaluations"]
        timings[mode] = ((time.perf_counter() - t_start) / args.frames * 1e3, ev, r)
    host_lm_ms = timings[False][0]
    gpu_ms, ev, r = timings[True]
    oc = orc.CoarseTracker(T["w"], T["h"], T["K"], args.levels)
    oc.make_coarse_depth(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    oc.set_new_frame(T["pyr_new"])
    ro = oc.track(R0, t0, 0.0, 0.0, precision=0)
    t_start = time.perf_counter()
    for _ in range(args.cpu_frames):
        ro = oc.track(R0, t0, 0.0, 0.0, precision=0)
    cpu_ms = (time.perf_counter() - t_start) / args.cpu_frames * 1e3
    pts_per_eval = float(np.mean(counts))
    out = {"metric": "coarse tracking ms/frame (trackNewestCoarse, 640x480)", "levels": L, "ref_points_per_level": counts,
           "gpu_ms_per_frame": gpu_ms, "gpu_ms_per_frame_host_lm_loop": host_lm_ms, "gpu_evaluations_per_frame": ev / args.frames, "gpu_us_per_evaluation": gpu_ms * 1e3 / (ev / args.frames),
           "gpu_points_per_s": sum(counts) / L * (ev / args.frames) / (gpu_ms * 1e-3),
           "set_ref_ms_device": set_ref_dev_ms, "set_ref_ms_host_build_plus_upload": set_ref_host_ms,
           "gpu_track_only_ms": track_only[True], "gpu_track_only_ms_host_lm_loop": track_only[False], "gpu_upload_pyramid_plus_one_eval_ms": upload_ms,
           "cpu_oracle_ms_per_frame": cpu_ms, "cpu_oracle_make_images_ms": cpu_pyr_ms, "cpu_threads": 1,
           "speedup_track_only": cpu_ms / track_only[True], "speedup_frame_incl_pyramid": (cpu_ms + cpu_pyr_ms) / gpu_ms, "speedup": cpu_ms / gpu_ms,
           "pose_error_vs_truth_t": float(np.linalg.norm(r["t"] - T["t_true"])), "gpu_vs_cpu_dt": float(np.abs(r["t"] - ro["t"]).max()),
           "iterations_gpu": r["iterations"], "iterations_cpu": ro["iterations"],
           "timed": "per frame: H2D of the raw image, device pyramid, the whole LM loop in ONE persistent launch (dmv_ct_track); host_lm_loop = 1 fused launch + sync per evaluation, 8x8 LDLT on the host"}
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
