"""Worker of tests/test_dropin.py.  run_case() drives oracle/ref.py (the REFERENCE's own objects behind the C harness) through the calls of the
hot path; which shared library oracle/ref.py loads is decided by DMV_REF_LIB *before* import:
  - oracle/_ref/libdso_ref.so           the unmodified reference
  - oracle/_ref/libdso_ref_dropin.so    the same reference objects with calcRes / calcGSSSE / accumulateAF_MT / accumulateSCF_MT /
                                         resubstituteF_MT replaced by oracle/dropin_stubs.cpp -> include/dmvio_b200.h
Run as a script it executes one case under the library named by DMV_REF_LIB and writes the results to an .npz (a second copy of the reference's
process-global state cannot live in the test process)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

BA_CASES = {
    "c1": dict(nf=4, npts=300, seed=7),
    "c3": dict(nf=7, npts=2000, seed=1234),   # BASELINE config 3
    "states": dict(nf=5, npts=500, seed=21),
}
CT_CASES = {
    "small": dict(w=160, h=120, seed=77, npts=400),
    "vga": dict(seed=4321),
}


def run_ba(ref, synth, name):
    cfg = BA_CASES[name]
    W = synth.make_window(**cfg)
    if name == "states":  # incoming OOB / OUTLIER states, old energies, depth priors
        rng = np.random.default_rng(cfg["seed"])
        n = len(W["res_point"])
        W["res_state"] = rng.choice([0, 1, 2], n, p=[0.8, 0.1, 0.1]).astype(np.int32)
        W["res_energy"] = rng.uniform(0, 50, n).astype(np.float32)
        W["hasDepthPrior"] = (rng.random(len(W["host"])) < 0.3).astype(np.uint8)
    rw = ref.Window(W)
    out = {}
    out["E"] = np.float64(rw.linearize_all(update_th=False))       # the reference's own PointFrameResidual::linearize on both sides
    rw.apply_res()
    a = rw.accumulate(0)                                              # accumulateAF_MT / LF / SCF_MT
    for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc"):
        out["acc_" + k] = a[k]
    out["resInA"] = np.int64(a["resInA"])
    p = rw.point_outputs()
    for k in ("Hdd", "bd", "Hcd", "HdiF", "bdSumF"):
        out["pt_" + k] = p[k]
    x, HS, bS = rw.solve(0, 1e-5, 0)                                  # EnergyFunctional::solveSystemF: accumulate*, dense solve, resubstituteF_MT
    out["x"], out["HS"], out["bS"] = x, HS, bS
    out["step"] = rw.point_outputs()["step"]
    out["E_hot"] = np.float64(rw.hot_iteration(x))                   # accumulate + resubstitute + step + linearizeAll + restore
    del rw
    return out


def run_ct(ref, synth, name):
    T = synth.make_tracking_pair(**CT_CASES[name])
    rc = ref.CoarseTracker(T["w"], T["h"], T["K"])
    rc.make_coarse_depth(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    rc.set_new_frame(T["pyr_new"], 1.0, 1.2, 0.01, -0.5)
    R, t = synth.se3_mul(*synth.se3_exp(np.array([0.003, -0.002, 0.001, 0.001, -0.001, 0.001])), T["R_true"], T["t_true"])
    a, b = T["a_new"] + 0.01, T["b_new"] - 0.3
    out = {"levels": np.int64(rc.levels)}
    for l in range(rc.levels):
        for ci, cutoff in enumerate((20.0, 5.0)):
            out[f"res_{l}_{ci}"] = rc.calc_res(l, R, t, a, b, cutoff)       # CoarseTracker::calcRes
            H, bb = rc.calc_gs(l, a, b, 0)                                 # CoarseTracker::calcGSSSE
            out[f"H_{l}_{ci}"], out[f"b_{l}_{ci}"] = H, bb
    tr = rc.track(np.eye(3), np.zeros(3), 0.0, 0.0, precision=0)           # CoarseTracker::trackNewestCoarse (the reference's LM loop)
    out["good"] = np.int64(tr["good"])
    out["R"], out["t"], out["ab"] = tr["R"], tr["t"], np.array([tr["a"], tr["b"]])
    out["lastResiduals"], out["flow"] = np.asarray(tr["lastResiduals"], np.float64), np.asarray(tr["flow"], np.float64)
    del rc
    return out


CI_CASES = {"few": (50, 45, 35, 25), "many": (3000, 900, 300, 100)}


def run_ci(ref, synth, name):
    """the reference's own CoarseInitializer::trackFrame (pyramid loop, LM, depth steps, regularisation, propagation, snapping) over six frames;
    in the drop-in build every calcResAndGS inside it is served by dmv_ci_calc_res_and_gs"""
    from helpers import init_points
    frames = [synth.make_tracking_pair(seed=77, trans=0.02 * k, rot=0.004 * k) for k in (1, 2, 3)]
    T = frames[0]
    w, h, L = T["w"], T["h"], T["levels"]
    pts = init_points(np.random.default_rng(6), w, h, L, CI_CASES[name])
    rc = ref.CoarseInit(w, h, T["K"])
    rc.set_first(T["pyr_ref"], 1.0, pts)
    rc.set_new(T["pyr_new"], 1.2)
    out = {"levels": np.int64(L)}
    R, t = synth.se3_exp(np.array([0.004, -0.003, 0.002, 0.002, -0.001, 0.0015]))
    for lvl in range(L):
        s = rc.calc(lvl, R, t, np.log(1.2), 0.3)                            # CoarseInitializer::calcResAndGS directly
        for k in ("H", "b", "Hsc", "bsc", "res"):
            out[f"calc_{lvl}_{k}"] = s[k]
        p = rc.points(lvl)
        out[f"calc_{lvl}_good"], out[f"calc_{lvl}_energy"] = p["isGood_new"], p["energy_new0"]
    rc.set_first(T["pyr_ref"], 1.0, pts)
    for k, F in enumerate(frames + frames[::-1]):
        tr = rc.track(F["pyr_new"], 1.0 + 0.05 * k)
        out[f"trk_{k}_state"] = np.array([tr["ok"], tr["snapped"], tr["snappedAt"], tr["frameID"]], np.int64)
        out[f"trk_{k}_R"], out[f"trk_{k}_t"], out[f"trk_{k}_ab"] = tr["R"], tr["t"], np.array([tr["a"], tr["b"]])
    p0 = rc.points(0)
    out["final_isGood"], out["final_idepth"] = p0["isGood"], p0["idepth"]
    del rc
    return out


def run_case(ref, synth, case):
    kind, name = case.split(":")
    return {"ba": run_ba, "ct": run_ct, "ci": run_ci}[kind](ref, synth, name)


if __name__ == "__main__":
    case, outfile = sys.argv[1], sys.argv[2]
    from oracle import orc as _orc
    _orc.build()
    from oracle import ref as _ref
    import dmvio_b200.synth as _synth
    res = run_case(_ref, _synth, case)
    np.savez(outfile, **res)
