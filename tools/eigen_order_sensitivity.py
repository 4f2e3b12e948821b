#!/usr/bin/env python
"""How much do the reference's results depend on the ASSOCIATION ORDER of small inner products inside the matrix library?

oracle/_ref/libdso_ref.so compiles the reference's own sources against a stand-in Eigen that sums inner products sequentially;
Eigen 3.3's unrolled reductions associate some sizes differently (halves splitting).  `oracle/ref_build.sh tree` builds the same
sources with the halves-splitting order.  This tool runs both builds on the same seeded inputs and reports the largest difference
of every pinned quantity -- an empirical bound on what "bit-exact against the reference" cannot cover without the real Eigen
headers.  TEST INFRASTRUCTURE: runs the compiled reference only, never the product.

    python tools/eigen_order_sensitivity.py            # prints a markdown table
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(out):
    import importlib
    from oracle import ref
    synth = importlib.import_module("dm-vio_b200.synth")
    res = {}
    W = synth.make_window(nf=7, npts=2000, seed=1234)
    rw = ref.Window(W)
    res["precalc"] = rw.precalc()
    res["E_lin"] = np.array([rw.linearize_all(update_th=False)])
    r = rw.res_outputs(True)
    res["res_newEnergy"] = r["newEnergy"]; res["res_newState"] = r["newState"]; res["res_J"] = r["J"]
    rw.apply_res()
    a = rw.accumulate(0)
    for k in ("HA", "bA", "Hsc", "bsc"):
        res[k] = a[k]
    p = rw.point_outputs()
    for k in ("Hdd", "bd", "HdiF"):
        res["pt_" + k] = p[k]
    x, _, _ = rw.solve(0, 1e-5, 0)
    res["x"] = x
    res["pt_step"] = rw.point_outputs()["step"]
    T = synth.make_tracking_pair(seed=4321)
    rc = ref.CoarseTracker(T["w"], T["h"], T["K"])
    rc.make_coarse_depth(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    rc.set_new_frame(T["pyr_new"], 1.0, 1.2, 0.01, -0.5)
    R, t = synth.se3_mul(*synth.se3_exp(np.array([0.003, -0.002, 0.001, 0.001, -0.001, 0.001])), T["R_true"], T["t_true"])
    res["ct_calcRes"] = rc.calc_res(0, R, t, T["a_new"], T["b_new"], 20.0)
    H, b = rc.calc_gs(0, T["a_new"], T["b_new"], 0)
    res["ct_H"], res["ct_b"] = H, b
    tr = rc.track(np.eye(3), np.zeros(3), 0.0, 0.0, precision=0)
    res["ct_track_t"] = tr["t"]; res["ct_track_R"] = tr["R"]; res["ct_track_res"] = np.asarray(tr["lastResiduals"][:rc.levels])
    import importlib as _il
    hm = _il.import_module("dm-vio_b200.hostmath")
    W = synth.make_window(nf=3, npts=10, seed=5, trans=0.05, rot=0.01)
    w, h = W["w"], W["h"]
    rng = np.random.default_rng(1)
    n = 3000
    u, v = rng.integers(10, w - 10, n), rng.integers(10, h - 10, n)
    st = ref.ip_init(W["dI"][0], w, h, W["K"], u, v)
    for i, new in enumerate((1, 2, 1)):
        KRKi, Kt, aff = hm.trace_tables(W, 0, new)
        st = ref.ip_trace(st, W["dI"][new], w, h, W["K"], KRKi, Kt, aff)
    for k in ("idepth_min", "idepth_max", "status", "lastTracePixelInterval", "quality"):
        if k in st:
            res["ip_" + k] = np.asarray(st[k])
    np.savez(out, **res)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        return child(sys.argv[2])
    libs = {"seq": os.path.join(ROOT, "oracle", "_ref", "libdso_ref.so"), "tree": os.path.join(ROOT, "oracle", "_ref", "libdso_ref_tree.so")}
    if not os.path.exists(libs["seq"]):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh")])
    if not os.path.exists(libs["tree"]):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh"), "tree"])
    outs = {}
    with tempfile.TemporaryDirectory() as d:
        for k, p in libs.items():
            f = os.path.join(d, k + ".npz")
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", f], env=dict(os.environ, DMV_REF_LIB=p))
            outs[k] = dict(np.load(f))
    print("| quantity | n | differing entries | entries off by > 1e-3 of max |value| | max abs diff | max diff / max |value| |")
    print("|---|---|---|---|---|---|")
    summary = {}
    for k in outs["seq"]:
        a, b = outs["seq"][k].astype(np.float64).ravel(), outs["tree"][k].astype(np.float64).ravel()
        ok = np.isfinite(a) & np.isfinite(b)
        nd = int((a != b).sum())
        md = float(np.abs(a[ok] - b[ok]).max()) if ok.any() else 0.0
        sc = float(np.abs(a[ok]).max()) if ok.any() else 1.0
        big = int((np.abs(a[ok] - b[ok]) > 1e-3 * sc).sum()) if ok.any() else 0
        summary[k] = dict(n=int(a.size), differing=nd, big=big, max_abs=md, rel=md / sc if sc else 0.0)
        print(f"| {k} | {a.size} | {nd} | {big} | {md:.3g} | {md / sc if sc else 0:.3g} |")
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
