"""Host logic of the C++ adapters (dm-vio_b200/host/window_ba.cpp, coarse_tracker.cpp) in the CPU test tier.

oracle/libhost_on_oracle.so = the adapter's UNMODIFIED sources linked against oracle/mock_capi.cpp, a CPU stand-in of the C ABI built on the
oracle (test infrastructure, never shipped).  The adapter therefore runs its own control flow here — FullSystem::optimize's LM loop with
solveSystemF (priors, marginalisation prior, Jacobi-preconditioned LDLT, gauge projection from iteration 2 on), backup / step / restore,
the tail (setEvalPT, linearizeAll(true) bookkeeping, residual deletion), flagPointsForRemoval, marginalizePointsF, marginalizeFrame — and is
compared with the oracle's own optimize / finishOptimize / marginalize, which tests/test_ref_pin.py pins to the reference.  Both sides use the
same residual arithmetic (the oracle's), so the tolerances are tight: what differs is only the adapter's host-side fp64 code and its float
tables.  The coarse-tracker adapter (makeK, host makeCoarseDepthL0, calcRes operand set-up, its own LM loop, the abort path) is driven the same
way against the oracle's trackNewestCoarse.  The same scenarios run against the CUDA kernels in tests/test_gpu_host.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libhost_on_oracle.so")


@pytest.fixture(scope="module")
def hostapi(orc):
    import dmvio_b200.hostapi as h
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libhost_on_oracle.so"])
    mock = h._bind(C.CDLL(LIB))
    saved = h._L
    h._L = mock            # WindowBA() picks the library up through hostapi.lib()
    yield h
    h._L = saved


def test_tables_and_first_linearisation(hostapi, orc, synth):
    W = synth.make_window(nf=5, npts=300, seed=3)
    ow = orc.Window(W)
    hw = hostapi.WindowBA(W)
    pc, adH, adT = hw.tables()
    np.testing.assert_allclose(pc, ow.precalc(), rtol=2e-6, atol=1e-4)
    a_o, t_o = ow.adjoints()
    np.testing.assert_allclose(adH, a_o, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(adT, t_o, rtol=1e-12, atol=1e-12)
    e_g, e_o = hw.linearize(), ow.linearize_all(update_th=False)
    assert abs(e_g - e_o) <= 1e-5 * abs(e_o)
    hw.close()


@pytest.mark.parametrize("cfg", [dict(nf=4, npts=600, seed=13), dict(nf=7, npts=800, seed=1234)], ids=["nf4", "nf7"])
def test_optimize_loop_matches_oracle(hostapi, orc, synth, cfg):
    """the LM loop incl. the gauge projection of x from iteration 2 on: same accept / reject sequence, energies and final states"""
    W = synth.make_window(state_noise=2e-3, **cfg)
    ow = orc.Window(W)
    n_o, log_o = ow.optimize(6, precision=1)
    hw = hostapi.WindowBA(W)
    n_g, log_g = hw.optimize(6)
    assert n_g == n_o and n_o >= 3                                   # at least one iteration with the projection active
    np.testing.assert_allclose(log_g, log_o, rtol=1e-5)
    st_g, id_g, th_g = hw.states()
    assert np.abs(st_g - ow.frame_states()).max() < 2e-6
    np.testing.assert_allclose(id_g, ow.point_outputs()["idepth"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(th_g, ow.frame_tables()["frameEnergyTH"], rtol=1e-4)
    # without the projection the oracle ends somewhere measurably different: the call site in WindowBA::solveSystemF matters
    ob = orc.Window(W, settings=dict(orthogonalizeXLater=0))
    ob.optimize(6, precision=1)
    assert np.abs(ob.frame_states() - ow.frame_states()).max() > 10 * np.abs(st_g - ow.frame_states()).max()
    hw.close()


def test_finish_optimize_and_second_optimize(hostapi, orc, synth):
    W = synth.make_window(nf=5, npts=600, seed=21, state_noise=2e-3)
    nres = len(W["res_point"])
    ow = orc.Window(W)
    ow.optimize(4, precision=1)
    E_o, rem_o = ow.finish_optimize()
    hw = hostapi.WindowBA(W)
    hw.optimize(4)
    E_g, rem_g = hw.finish_optimize()
    assert abs(E_g - E_o) <= 1e-5 * abs(E_o)
    np.testing.assert_array_equal(rem_g, rem_o)
    assert hw.nres == nres - len(rem_g)
    ps_o, ps_g = ow.point_stats(), hw.point_stats()
    np.testing.assert_array_equal(ps_g["numGoodResiduals"], ps_o["numGoodResiduals"])
    np.testing.assert_allclose(ps_g["maxRelBaseline"], ps_o["maxRelBaseline"], rtol=1e-4, atol=1e-7)
    st_g, _, th_g = hw.states()
    assert np.all(st_g[-1, :6] == 0)
    assert np.abs(st_g - ow.frame_states()).max() < 2e-6
    n_o, log_o = ow.optimize(3, precision=1)
    n_g, log_g = hw.optimize(3)                                       # resetOOB on the thinned window; deleted residuals stay out
    assert n_g == n_o
    np.testing.assert_allclose(log_g, log_o, rtol=1e-5)
    hw.close()


def test_point_marginalisation_matches_oracle(hostapi, orc, synth):
    W = synth.make_window(nf=5, npts=800, seed=17, state_noise=2e-3)
    ow = orc.Window(W)
    ow.optimize(4, precision=1)
    hw = hostapi.WindowBA(W)
    hw.optimize(4)
    po = ow.point_outputs()
    idepth_hessian = np.where(po["HdiF"] > 0, 1.0 / np.maximum(po["HdiF"], 1e-30), 0.0)
    rng = np.random.default_rng(0)
    well = np.nonzero(idepth_hessian > 200)[0]
    marg = np.sort(rng.choice(well, len(well) // 3, replace=False)).astype(np.int32)
    weak = np.nonzero((idepth_hessian > 0) & (idepth_hessian < 40))[0][:5].astype(np.int32)     # below setting_minIdepthH_marg: dropped, not marginalised
    rest = np.setdiff1d(np.arange(len(W["host"])), np.concatenate([marg, weak]))
    drop = np.sort(rng.choice(rest, 20, replace=False)).astype(np.int32)
    o = ow.marginalize(marg, precision=1)
    g = hw.marginalize_points(np.concatenate([marg, weak]), drop)
    assert g["npts"] == len(W["host"]) - len(marg) - len(weak) - len(drop)
    assert g["nres"] == int((~np.isin(W["res_point"], np.concatenate([marg, weak, drop]))).sum())
    assert g["resInM"] == o["resInM"]
    assert rel(g["HM"], o["HM"]) < 1e-5 and rel(g["bM"], o["bM"]) < 1e-4
    e = hw.linearize()
    assert np.isfinite(e) and e > 0
    hw.close()


def test_keyframe_turnover_flow(hostapi, orc, synth):
    nf = 6
    W = synth.make_window(nf=nf, npts=500, seed=31, state_noise=1e-3, hosts="all")
    prior0 = orc.Window(W).frame_tables()["prior"][0]
    hw = hostapi.WindowBA(W)
    hw.optimize(3)
    E, rem = hw.finish_optimize()
    res_point = np.asarray(W["res_point"])[np.setdiff1d(np.arange(len(W["res_point"])), rem)]
    res_target = np.asarray(W["res_target"])[np.setdiff1d(np.arange(len(W["res_target"])), rem)]
    for _ in range(3):
        assert len(hw.finish_optimize()[1]) == 0                       # nothing left to delete; numGoodResiduals keeps counting
    npts = len(W["host"])
    rng = np.random.default_rng(4)
    last_t = np.tile(np.asarray(W["frameID"])[[-1, -2]], (npts, 1)).astype(np.int32)
    last_s = rng.choice([0, 1, 2], (npts, 2), p=[0.8, 0.1, 0.1]).astype(np.int32)
    hw.set_last_residuals(last_t, last_s)
    marg, drop = hw.flag_points([0])
    # ---- flagPointsForRemoval against a numpy restatement of PointHessian::isOOB / isInlierNew (HessianBlocks.h:L476-506)
    _, idepth, _ = hw.states()
    ps = hw.point_stats()
    nres_p = np.bincount(res_point, minlength=npts)
    vis = np.bincount(res_point[res_target == 0], minlength=npts)
    exp_m, exp_d = [], []
    for i in range(npts):
        if idepth[i] < 0.02 or nres_p[i] == 0:
            exp_d.append(i); continue
        oob = (nres_p[i] >= 3 and ps["numGoodResiduals"][i] > 14 and nres_p[i] - vis[i] < 3)
        if not oob:
            oob = last_s[i, 0] == 1 or (nres_p[i] >= 2 and last_s[i, 0] == 2 and last_s[i, 1] == 2)
        if not oob and W["host"][i] != 0:
            continue
        (exp_m if (nres_p[i] >= 3 and ps["numGoodResiduals"][i] >= 4) else exp_d).append(i)
    np.testing.assert_array_equal(marg, np.asarray(exp_m, np.int32))
    np.testing.assert_array_equal(drop, np.asarray(exp_d, np.int32))
    assert len(marg) > 20 and len(drop) > 0
    # ---- marginalise the points, then the frame
    g = hw.marginalize_points(marg, drop)
    assert g["resInM"] > 0
    st, _, _ = hw.states()
    exp_H, exp_b = hostapi.marginalize_frame_hm(g["HM"], g["bM"], nf, 0, prior0, st[0][:8])
    nres_before = hw.nres
    m = hw.marginalize_frame(0)
    assert m["nf"] == nf - 1 and rel(m["HM"], exp_H) < 1e-12 and rel(m["bM"], exp_b) < 1e-12
    assert m["nres"] == nres_before - int(np.isin(res_point, np.setdiff1d(np.arange(npts), np.concatenate([marg, drop])))[res_target == 0].sum())
    e = hw.linearize()
    eL, eM = hw.energies_LM()
    assert np.isfinite(e) and e > 0 and eM != 0
    n, log = hw.optimize(3)
    eL1, eM1 = hw.energies_LM()
    assert n >= 1 and np.all(np.isfinite(log))
    assert log[-1] + eL1 + eM1 <= (log[0] + eL + eM) * (1 + 1e-9)
    hw.close()


@pytest.mark.parametrize("device_lm", [False, True], ids=["host_lm_loop", "one_call_track"])
@pytest.mark.parametrize("levels", [4, 5])
def test_coarse_tracker_adapter(hostapi, orc, synth, levels, device_lm):
    """host/coarse_tracker.cpp on the stand-in C ABI: makeK, the host-side makeCoarseDepthL0 + reference upload, the operand set-up of
    calcRes (R*Ki in float with the reference's K inverse, affLL) and — with device_lm = False — the adapter's own Levenberg-Marquardt loop of
    trackNewestCoarse (level repeats, cutoff doubling, accept / reject, 8x8 solve) against the oracle's trackNewestCoarse."""
    T = synth.make_tracking_pair(seed=4321, levels=levels if levels == 5 else 0)
    oct_ = orc.CoarseTracker(T["w"], T["h"], T["K"], levels if levels == 5 else 0)
    oct_.make_coarse_depth(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    oct_.set_new_frame(T["pyr_new"])
    r_o = oct_.track(np.eye(3), np.zeros(3), 0.0, 0.0)
    g = hostapi.CoarseTracker(T["w"], T["h"], T["K"], levels)
    counts = g.set_ref(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    assert counts == [len(oct_.ref_points(l)["u"]) for l in range(levels)]
    g.set_new_image(T["img_new"])
    r_g = g.track(np.eye(3), np.zeros(3), 0.0, 0.0, device_lm=device_lm)
    assert r_g["good"] == r_o["good"]
    assert np.abs(r_g["R"] - r_o["R"]).max() < 1e-7 and np.abs(r_g["t"] - r_o["t"]).max() < 1e-7
    assert abs(r_g["a"] - r_o["a"]) < 1e-6 and abs(r_g["b"] - r_o["b"]) < 1e-4
    np.testing.assert_allclose(r_g["lastResiduals"][:levels], r_o["lastResiduals"][:levels], rtol=1e-6)
    assert r_g["iterations"] == r_o["iterations"]
    assert np.linalg.norm(r_g["t"] - T["t_true"]) < 5e-4
    # setCoarseTrackingRef with makeCoarseDepthL0 behind the C ABI (raw keyframe image in) gives the same lists and track
    b = hostapi.CoarseTracker(T["w"], T["h"], T["K"], levels)
    assert b.set_ref_device(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["img_ref"]) == counts
    b.set_new_image(T["img_new"])
    r_b = b.track(np.eye(3), np.zeros(3), 0.0, 0.0, device_lm=device_lm)
    np.testing.assert_array_equal(r_b["R"], r_g["R"]); np.testing.assert_array_equal(r_b["t"], r_g["t"])
    g.close(); b.close()


def test_coarse_tracker_abort_leaves_outputs_untouched(hostapi, orc, synth):
    """residual above 1.5 x minResForAbort: trackNewestCoarse returns false before it writes its outputs (CoarseTracker.cpp:L731-735), on both
    paths of the adapter"""
    T = synth.make_tracking_pair(seed=4321)
    L = T["levels"]
    R0, t0 = synth.se3_exp(np.array([0.01, 0.0, 0.0, 0.0, 0.002, 0.0]))
    for device_lm in (False, True):
        g = hostapi.CoarseTracker(T["w"], T["h"], T["K"], L)
        g.set_ref(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
        g.set_new_image(T["img_new"])
        r = g.track(R0, t0, 0.1, 0.2, minRes=np.full(5, 1e-6), device_lm=device_lm)
        assert not r["good"]
        np.testing.assert_array_equal(r["R"], R0); np.testing.assert_array_equal(r["t"], t0)
        assert r["a"] == 0.1 and r["b"] == 0.2
        g.close()


def _subwindow(W, frames, pts_mask=None):
    """the window restricted to `frames` (sorted list of frame indices of W): points hosted there, residuals with both ends inside"""
    frames = list(frames)
    fmap = {f: i for i, f in enumerate(frames)}
    host = np.asarray(W["host"])
    pm = np.isin(host, frames) if pts_mask is None else pts_mask
    pidx = np.nonzero(pm)[0]
    pmap = -np.ones(len(host), np.int64); pmap[pidx] = np.arange(len(pidx))
    rp, rt = np.asarray(W["res_point"]), np.asarray(W["res_target"])
    rm = pm[rp] & np.isin(rt, frames)
    S = dict(W)
    S.update(nf=len(frames), dI=[W["dI"][f] for f in frames], images=[W["images"][f] for f in frames], R_eval=W["R_eval"][frames], t_eval=W["t_eval"][frames],
             state=W["state"][frames], state_zero=W["state_zero"][frames], exposure=W["exposure"][frames], frameEnergyTH=W["frameEnergyTH"][frames],
             frameID=W["frameID"][frames], host=np.array([fmap[h] for h in host[pidx]], np.int32))
    for k in ("u", "v", "idepth", "idepth_zero", "color", "weights", "hasDepthPrior"):
        S[k] = np.asarray(W[k])[pidx]
    S["res_point"] = pmap[rp[rm]].astype(np.int32)
    S["res_target"] = np.array([fmap[t] for t in rt[rm]], np.int32)
    return S, pidx, np.nonzero(rm)[0]


def test_keyframe_stream_with_prior_matches_oracle(hostapi, orc, synth):
    """One full keyframe turnover followed by the next optimisation, against the oracle: optimize -> tail -> flag / marginalise the points of the
    oldest keyframe -> marginalise the keyframe (prior Schur-complemented, indices shift, image slot freed) -> a NEW keyframe enters (prior grows
    by 8 zero rows / columns, the free slot is reused) with its points and residuals -> optimize.  The oracle gets the same window built from
    scratch with the adapter's prior; energies and poses must agree, i.e. the adapter uses HM / bM exactly as EnergyFunctional does."""
    W = synth.make_window(nf=6, npts=600, seed=37, state_noise=1e-3, hosts="all")
    A, pidxA, ridxA = _subwindow(W, [0, 1, 2, 3, 4])
    hw = hostapi.WindowBA(A)
    hw.optimize(3)
    st_pre, _, _ = hw.states()
    E, rem = hw.finish_optimize()
    marg, drop = hw.flag_points([0])
    g = hw.marginalize_points(marg, drop)
    m = hw.marginalize_frame(0)
    st, idepth, th = hw.states()
    # ---- what the adapter holds now, tracked on the Python side (orders are preserved by every erase)
    keep_r = np.ones(len(A["res_point"]), bool); keep_r[rem] = False
    gone_p = np.zeros(len(A["host"]), bool); gone_p[marg] = True; gone_p[drop] = True
    keep_r &= ~gone_p[A["res_point"]] & (A["res_target"] != 0)
    kp = np.nonzero(~gone_p)[0]
    assert len(kp) == hw.npts and int(keep_r.sum()) == hw.nres
    pmap = -np.ones(len(A["host"]), np.int64); pmap[kp] = np.arange(len(kp))
    # ---- the new keyframe (frame 5 of W) with the points it hosts and every residual between it and the kept frames
    new_p = np.nonzero(np.asarray(W["host"]) == 5)[0]
    host2 = np.concatenate([A["host"][kp] - 1, np.full(len(new_p), 4)]).astype(np.int32)
    cat = lambda k: np.concatenate([np.asarray(A[k])[kp], np.asarray(W[k])[new_p]])
    idepth2 = np.concatenate([idepth, np.asarray(W["idepth"])[new_p]]).astype(np.float32)
    rp, rt = np.asarray(W["res_point"]), np.asarray(W["res_target"])
    gmap = -np.ones(len(W["host"]), np.int64)                      # W point index -> new window point index
    gmap[pidxA[kp]] = np.arange(len(kp)); gmap[new_p] = len(kp) + np.arange(len(new_p))
    newr = (gmap[rp] >= 0) & (((rt == 5) & (np.asarray(W["host"])[rp] != 5) & (np.asarray(W["host"])[rp] >= 1)) | ((np.asarray(W["host"])[rp] == 5) & (rt >= 1)))
    res_point2 = np.concatenate([pmap[A["res_point"][keep_r]], gmap[rp[newr]]]).astype(np.int32)
    res_target2 = np.concatenate([A["res_target"][keep_r] - 1, rt[newr] - 1]).astype(np.int32)
    L = hw.L
    c = lambda a, t: np.ascontiguousarray(a, t)
    idx = L.dmvh_window_add_frame(hw.h, c(W["dI"][5], np.float32).reshape(-1), 0, c(W["R_eval"][5], np.float64).reshape(-1), c(W["t_eval"][5], np.float64),
                                  c(W["state"][5], np.float64), c(W["state_zero"][5], np.float64), float(W["exposure"][5]), int(W["frameID"][5]))
    assert idx == 4
    L.dmvh_window_set_points(hw.h, len(host2), host2, c(cat("u"), np.float32), c(cat("v"), np.float32), idepth2, idepth2, c(cat("color"), np.float32).reshape(-1),
                             c(cat("weights"), np.float32).reshape(-1), None)
    L.dmvh_window_set_residuals(hw.h, len(res_point2), res_point2, res_target2)
    assert L.dmvh_window_prepare(hw.h) == 0
    hw.nf, hw.N, hw.npts, hw.nres = 5, 44, len(host2), len(res_point2)
    # ---- the same window for the oracle
    R4, t4 = synth.se3_mul(*synth.se3_exp(st_pre[4][:6]), A["R_eval"][4], A["t_eval"][4])       # setEvalPT moved frame 4's evaluation point onto its estimate
    HM = np.zeros((44, 44)); HM[:36, :36] = m["HM"]
    bM = np.zeros(44); bM[:36] = m["bM"]
    st2 = np.concatenate([st, W["state"][5][None]], 0)
    sz2 = np.concatenate([A["state_zero"][1:4], np.r_[np.zeros(6), st[3][6:8], 0, 0][None], W["state_zero"][5][None]], 0)
    W2 = dict(w=W["w"], h=W["h"], nf=5, K=W["K"], dI=[W["dI"][f] for f in (1, 2, 3, 4, 5)],
              R_eval=np.concatenate([A["R_eval"][1:4], R4[None], W["R_eval"][5][None]], 0), t_eval=np.concatenate([A["t_eval"][1:4], t4[None], W["t_eval"][5][None]], 0),
              state=st2, state_zero=sz2, exposure=np.ones(5, np.float32), frameEnergyTH=np.r_[th, 512.0].astype(np.float32), frameID=np.arange(1, 6, dtype=np.int32),
              host=host2, u=cat("u"), v=cat("v"), idepth=idepth2, idepth_zero=idepth2, color=cat("color"), weights=cat("weights"),
              hasDepthPrior=np.zeros(len(host2), np.uint8), res_point=res_point2, res_target=res_target2, HM=HM, bM=bM)
    ow = orc.Window(W2)
    n_o, log_o = ow.optimize(4, precision=1)
    n_g, log_g = hw.optimize(4)
    assert n_g == n_o
    np.testing.assert_allclose(log_g, log_o, rtol=5e-5)              # the adapter's float tables differ from the oracle's in the last bits (see test_tables_*)
    st_g, _, _ = hw.states()
    assert np.abs(st_g - ow.frame_states()).max() < 5e-6
    # the prior matters: without it the oracle lands somewhere else
    W3 = dict(W2); W3.pop("HM"); W3.pop("bM")
    o3 = orc.Window(W3); o3.optimize(4, precision=1)
    assert np.abs(o3.frame_states() - ow.frame_states()).max() > 10 * np.abs(st_g - ow.frame_states()).max()
    hw.close()


def test_insert_points_on_older_host_keeps_the_other_points_statistics(hostapi, synth):
    """newly activated points are hosted by OLDER keyframes, so they land in the middle of the host-ordered list and every later
    point changes index.  insertPoints resets the per-point statistics (numGoodResiduals, maxRelBaseline, lastResiduals) unless carry_from maps
    an entry to its previous index; flagPointsForRemoval (isOOB / isInlierNew) must then see each point's OWN history."""
    W = synth.make_window(nf=4, npts=300, seed=11)
    hw = hostapi.WindowBA(W)
    hw.optimize(3)
    hw.finish_optimize()                       # linearizeAll(true): numGoodResiduals / maxRelBaseline are now non-trivial
    s0 = hw.point_stats()
    assert s0["numGoodResiduals"].max() > 0 and len(set(np.round(s0["maxRelBaseline"], 6))) > 10
    _, idepth, _ = hw.states()
    n = hw.npts
    host = np.asarray(W["host"])
    # 20 new points hosted by frame 0 (the oldest): inserted after frame 0's existing points, i.e. in the middle of the list
    n0 = int((host == 0).sum())
    new = np.arange(20)
    order_old = np.concatenate([np.arange(n0), -np.ones(20, np.int64), np.arange(n0, n)])   # previous index of every entry, -1 = new
    pick = lambda a, fill: np.concatenate([np.asarray(a)[:n0], fill, np.asarray(a)[n0:]])
    host2 = pick(host, np.zeros(20, np.int32)).astype(np.int32)
    u2, v2 = pick(W["u"], W["u"][new] + 3).astype(np.float32), pick(W["v"], W["v"][new] + 3).astype(np.float32)
    id2 = pick(idepth, idepth[new]).astype(np.float32)
    col2 = np.concatenate([W["color"][:n0], W["color"][new], W["color"][n0:]]).astype(np.float32)
    wgt2 = np.concatenate([W["weights"][:n0], W["weights"][new], W["weights"][n0:]]).astype(np.float32)
    c = lambda a, t: np.ascontiguousarray(a, t)
    L = hw.L
    L.dmvh_window_set_points_carry(hw.h, len(host2), host2, u2, v2, id2, id2, c(col2, np.float32).reshape(-1), c(wgt2, np.float32).reshape(-1), None,
                                   c(order_old, np.int32))
    hw.npts = len(host2)
    s1 = hw.point_stats()
    old = order_old >= 0
    np.testing.assert_array_equal(s1["numGoodResiduals"][old], s0["numGoodResiduals"])      # every surviving point kept ITS OWN history
    np.testing.assert_array_equal(s1["maxRelBaseline"][old], s0["maxRelBaseline"])
    assert np.all(s1["numGoodResiduals"][~old] == 0) and np.all(s1["maxRelBaseline"][~old] == 0)
    # without the map everything is reset (never another point's values)
    L.dmvh_window_set_points(hw.h, len(host2), host2, u2, v2, id2, id2, c(col2, np.float32).reshape(-1), c(wgt2, np.float32).reshape(-1), None)
    s2 = hw.point_stats()
    assert np.all(s2["numGoodResiduals"] == 0) and np.all(s2["maxRelBaseline"] == 0)
    hw.close()


def test_compute_ba_update_hook_receives_the_reference_conventions(hostapi, orc, synth):
    """SURVEY §8b consumer contract: WindowBA::computeBAUpdate is the slot of BAGTSAMIntegration::computeBAUpdate(H, b, lambda, frames, HNoLambda) -> x.
    A hook that solves H x = b like the no-GTSAM branch (Jacobi-preconditioned) must reproduce the built-in path; H must be the lambda-damped
    Schur-reduced system in DSO ordering, HNoLambda the undamped one."""
    import ctypes as C
    W = synth.make_window(nf=4, npts=300, seed=19)
    ref = hostapi.WindowBA(W); n_ref, log_ref = ref.optimize(4); st_ref, _, _ = ref.states(); ref.close()
    calls = []

    CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

    def cb(H, b, lam, nframes, H0, x, user):
        N = 8 * nframes + 4
        Hm = np.ctypeslib.as_array(H, (N, N)).copy(); bm = np.ctypeslib.as_array(b, (N,)).copy(); H0m = np.ctypeslib.as_array(H0, (N, N)).copy()
        calls.append((lam, N, np.abs(Hm - Hm.T).max() / np.abs(Hm).max(), np.linalg.eigvalsh(0.5 * (H0m + H0m.T)).min()))
        sv = 1.0 / np.sqrt(np.diag(Hm) + 10)
        xs = sv * np.linalg.solve(sv[:, None] * Hm * sv[None, :], sv * bm)
        for i in range(N):
            x[i] = xs[i]

    hook = CB(cb)
    hw = hostapi.WindowBA(W)
    hw.L.dmvh_window_set_ba_update_hook.argtypes = [C.c_void_p, CB, C.c_void_p]
    hw.L.dmvh_window_set_ba_update_hook(hw.h, hook, None)
    n_h, log_h = hw.optimize(4)
    st_h, _, _ = hw.states()
    hw.close()
    assert n_h == n_ref and len(calls) == n_h
    assert all(c[1] == 8 * 4 + 4 and c[2] < 1e-9 for c in calls)          # N, symmetric
    assert calls[0][0] == pytest.approx(1e-5)                               # first lambda
    np.testing.assert_allclose(log_h, log_ref, rtol=2e-5)   # numpy's LU vs the adapter's pivoted LDL^T on a system conditioned ~1e7
    assert np.abs(st_h - st_ref).max() < 2e-6


def test_coarse_initializer_adapter_matches_oracle(hostapi, orc, synth):
    """dmvio_b200::CoarseInitializer::trackFrame (host loop over the pyramid, LM control, doStep / applyStep / optReg / propagateUp / Down /
    resetPoints) on the CPU stand-in of dmv_ci_calc_res_and_gs vs the oracle's trackFrame (pinned to the reference's compiled code): same
    decisions over six frames, poses and depths to float accumulation order."""
    from helpers import init_points
    frames = [synth.make_tracking_pair(seed=77, trans=0.02 * k, rot=0.004 * k) for k in (1, 2, 3)]
    T = frames[0]
    w, h, L = T["w"], T["h"], T["levels"]
    pts = init_points(np.random.default_rng(6), w, h, L, (3000, 900, 300, 100))
    oc = orc.CoarseInit(w, h, T["K"])
    assert oc.levels == L
    oc.set_first(T["pyr_ref"], 1.0, pts)
    g = hostapi.CoarseInit(w, h, T["K"], L, max_points=3008)
    g.set_first(T["pyr_ref"], 1.0, pts)
    for k, F in enumerate(frames + frames[::-1]):
        to, tg = oc.track(F["pyr_new"], 1.0 + 0.05 * k), g.track(F["pyr_new"], 1.0 + 0.05 * k)
        assert (tg["ok"], tg["snapped"], tg["snappedAt"], tg["frameID"]) == (to["ok"], to["snapped"], to["snappedAt"], to["frameID"]), k
        np.testing.assert_allclose(tg["R"], to["R"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(tg["t"], to["t"], rtol=0, atol=2e-4 * max(1e-2, np.abs(to["t"]).max()))
        assert abs(tg["a"] - to["a"]) < 1e-6 and abs(tg["b"] - to["b"]) < 1e-4
        po, pg = oc.points(0), g.points(0)
        same = po["isGood"] == pg["isGood"]
        assert same.mean() > 0.995
        assert np.median(np.abs(pg["idepth"][same] - po["idepth"][same]) / np.abs(po["idepth"][same])) < 1e-4
    assert tg["snapped"] and tg["evaluations"] > 50
    g.close()
