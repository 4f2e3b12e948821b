"""CPU tests pinning the oracle's BA restatement by first principles (the reference has no tests on this path):
finite differences of every Jacobian block, closed-form zero-residual case, structural invariants, fp32-faithful vs
fp64 accumulation, threaded vs single-threaded reduce, and GN convergence on a consistent synthetic scene."""
import numpy as np
import pytest

from oracle.orc import J_JABF, J_JIDX, J_JPDC, J_JPDD, J_JPDXI, J_RESF


def _fej_window(synth, ramp=True, nf=3, npts=60, seed=3):
    W = synth.make_window(nf=nf, npts=npts, seed=seed)
    h, w = W["h"], W["w"]
    if ramp:
        ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
        for k in range(nf):
            img = (0.3 + 0.05 * k) * xs + (0.2 - 0.03 * k) * ys + 40 + 5 * k
            W["dI"][k] = synth.make_pyramid(img.astype(np.float32), 1)[0].reshape(-1).copy()
    W["state"][:, :6] = 0
    W["state_zero"] = W["state"].copy()
    W["idepth_zero"] = W["idepth"].copy()
    return W


def _abs_jacobian(ow, o, ri, host, target, nf):
    """analytic d resF / d state_host, d state_target (8 each), d calib (4), d idepth for all 8 pixels"""
    J = o["J"][ri]
    JIdx = J[J_JIDX:J_JIDX + 16].reshape(2, 8)
    Jpdxi = J[J_JPDXI:J_JPDXI + 12].reshape(2, 6)
    Jpdc = J[J_JPDC:J_JPDC + 8].reshape(2, 4)
    Jpdd = J[J_JPDD:J_JPDD + 2]
    JabF = J[J_JABF:J_JABF + 16].reshape(2, 8)
    Jrel = np.zeros((8, 8))
    Jrel[:, :6] = JIdx.T @ Jpdxi
    Jrel[:, 6] = JabF[0]
    Jrel[:, 7] = JabF[1]
    adH, adT = ow.adjoints()
    idx = host + target * nf
    return Jrel @ adH[idx].T, Jrel @ adT[idx].T, JIdx.T @ Jpdc, JIdx.T @ Jpdd, JabF[1]


def test_jacobians_finite_difference(orc, synth):
    W = _fej_window(synth)
    ow = orc.Window(W)
    ow.linearize_all(update_th=False)
    o = ow.res_outputs()
    nf = W["nf"]
    checked = 0
    for ri in range(ow.nres):
        if o["newState"][ri] == 1:
            continue
        host, target = W["host"][W["res_point"][ri]], W["res_target"][ri]
        Jh, Jt, Jc, Jd, hw = _abs_jacobian(ow, o, ri, host, target, nf)
        gmax = np.abs(o["J"][ri][J_JIDX:J_JIDX + 16]).max()
        for which, Jan, eps in (("h", Jh, 1e-6), ("t", Jt, 1e-6)):
            for k in range(8):
                d = np.zeros(8)
                e = eps * (1.0 if k < 6 else (1e2 if k == 6 else 1.0))
                d[k] = e
                rp = ow.eval_raw(ri, dsh=d if which == "h" else None, dst=d if which == "t" else None)
                d[k] = -e
                rm = ow.eval_raw(ri, dsh=d if which == "h" else None, dst=d if which == "t" else None)
                if rp is None or rm is None:
                    continue
                fd = hw * (rp - rm) / (2 * e)
                scale = np.abs(Jan[:, k]).max() + 1e-3
                # centre pixel (pattern index 4) is exact; the others share the centre's geometric Jacobian
                assert abs(fd[4] - Jan[4, k]) < 2e-3 * scale + 1e-4, (ri, which, k, fd[4], Jan[4, k])
                # (DSO's approximation): the error is bounded by |grad| * pattern radius (2 px) for rotations
                assert np.abs(fd - Jan[:, k]).max() < 0.05 * scale + 2.5 * gmax + 1e-3, (ri, which, k)
        for k in range(4):
            d = np.zeros(4); e = 1e-5; d[k] = e
            rp = ow.eval_raw(ri, dcalib=d); rm = ow.eval_raw(ri, dcalib=-d)
            if rp is None or rm is None:
                continue
            fd = hw * (rp - rm) / (2 * e)
            scale = np.abs(Jc[:, k]).max() + 1e-3
            assert abs(fd[4] - Jc[4, k]) < 2e-3 * scale + 1e-3, (ri, "calib", k, fd[4], Jc[4, k])
        e = 1e-5
        rp = ow.eval_raw(ri, didepth=e); rm = ow.eval_raw(ri, didepth=-e)
        if rp is not None and rm is not None:
            fd = hw * (rp - rm) / (2 * e)
            assert abs(fd[4] - Jd[4]) < 2e-3 * (abs(Jd[4]) + 1e-3) + 1e-3
        checked += 1
    assert checked > 50


def test_residual_matches_first_principles(orc, synth):
    W = _fej_window(synth, ramp=False)
    ow = orc.Window(W)
    ow.linearize_all(update_th=False)
    o = ow.res_outputs()
    n = 0
    for ri in range(ow.nres):
        if o["newState"][ri] == 1:
            continue
        raw = ow.eval_raw(ri)
        J = o["J"][ri]
        hw = J[J_JABF + 8:J_JABF + 16]
        np.testing.assert_allclose(J[J_RESF:J_RESF + 8], hw * raw, rtol=2e-3, atol=2e-3)
        n += 1
    assert n > 50


def test_zero_residual_identity(orc, synth):
    """identical images + identity relative pose + neutral affine => r = 0, E = 0."""
    W = synth.make_window(nf=2, npts=50, seed=5)
    W["dI"][1] = W["dI"][0].copy()
    W["R_eval"][:] = np.eye(3); W["t_eval"][:] = 0
    W["state"][:] = 0; W["state_zero"][:] = 0
    ow = orc.Window(W)
    E = ow.linearize_all(update_th=False)
    o = ow.res_outputs()
    assert abs(E) < 1e-3
    assert np.all(o["newState"] == 0)
    # K * I * K.inverse() is the identity only up to the rounding of the reference's float cofactor inverse (~1e-5 px at x ~ 600),
    # times image gradients of up to ~30 / px
    assert np.abs(o["J"][:, J_RESF:J_RESF + 8]).max() < 5e-3


def test_structure_and_precision(orc, synth):
    W = synth.make_window(nf=4, npts=400, seed=7)
    ow = orc.Window(W)
    ow.linearize_all()
    ow.apply_res()
    a64 = ow.accumulate(1)
    a32 = ow.accumulate(0)
    for k in ("HA", "Hsc"):
        H = a64[k]
        assert np.abs(H - H.T).max() <= 1e-6 * np.abs(H).max()
        assert np.linalg.norm(a32[k] - H) / np.linalg.norm(H) < 2e-6
    for k in ("bA", "bsc"):
        assert np.linalg.norm(a32[k] - a64[k]) / np.linalg.norm(a64[k]) < 2e-5
    # Schur-reduced system (with priors) is PSD up to rounding
    S = a64["HA"] + a64["HL"] - a64["Hsc"]
    ev = np.linalg.eigvalsh(S)
    assert ev.min() > -1e-9 * ev.max()
    assert a64["resInA"] == int(ow.res_outputs(False)["isActive"].sum())


def test_threaded_reduce_matches(orc, synth):
    W = synth.make_window(nf=4, npts=500, seed=11)
    o1 = orc.Window(W, nthreads=1)
    o6 = orc.Window(W, nthreads=6)
    e1, e6 = o1.linearize_all(), o6.linearize_all()
    assert abs(e1 - e6) < 1e-6 * abs(e1)
    o1.apply_res(); o6.apply_res()
    a1, a6 = o1.accumulate(0), o6.accumulate(0)
    for k in ("HA", "bA", "Hsc", "bsc"):
        assert np.linalg.norm(a1[k] - a6[k]) / np.linalg.norm(a1[k]) < 1e-5
    x1, _, _ = o1.solve(); x6, _, _ = o6.solve()
    np.testing.assert_allclose(o1.point_outputs()["step"], o6.point_outputs()["step"], rtol=1e-3, atol=1e-5)


def test_gn_converges(orc, synth):
    W = synth.make_window(nf=4, npts=400, seed=13, state_noise=2e-3)
    ow = orc.Window(W)
    n, log = ow.optimize(6)
    assert n >= 1
    assert log[-1] < 0.5 * log[0]
    assert np.all(np.diff(log) <= 1e-9 * log[0])


def test_finish_optimize_tail(orc, synth):
    """tail of FullSystem::optimize (FullSystemOptimize.cpp:L591-609): the newest frame is re-based on its estimate without moving
    (precalc tables unchanged up to rounding), linearizeAll(true) counts the active residuals and lists the others for deletion."""
    W = synth.make_window(nf=5, npts=600, seed=21, state_noise=2e-3)
    ow = orc.Window(W)
    ow.optimize(4)
    pc0 = ow.precalc()
    st0 = ow.frame_states()
    E, removed = ow.finish_optimize()
    st1 = ow.frame_states()
    assert np.all(st1[-1, :6] == 0) and np.all(st1[-1, 6:8] == st0[-1, 6:8])      # setEvalPT: pose part of the state folded into the evaluation point
    np.testing.assert_array_equal(st1[:-1], st0[:-1])
    pc1 = ow.precalc().reshape(-1, 32); pc0 = pc0.reshape(-1, 32)
    cur = np.r_[0:12, 24:26]                                                       # current-state tables: KRKi, Kt, aff (the *_0 / b0 entries follow the new FEJ point)
    np.testing.assert_allclose(pc1[:, cur], pc0[:, cur], rtol=1e-5, atol=1e-4)     # nothing moved: exp(0) * (exp(xi) * T) == exp(xi) * T
    assert np.isfinite(E) and E > 0
    r = ow.res_outputs(False)
    keep = np.ones(ow.nres, bool); keep[removed] = False
    assert np.all(r["state"][removed] != 0) and np.all(r["state"][keep] == 0)     # exactly the non-IN residuals leave
    assert np.all(r["isActive"][keep] == 1)
    ps = ow.point_stats()
    good = np.bincount(np.asarray(W["res_point"])[keep], minlength=ow.npts)
    np.testing.assert_array_equal(ps["numGoodResiduals"], good)
    assert np.all(ps["maxRelBaseline"][good > 0] > 0) and np.all(ps["maxRelBaseline"][good == 0] == 0)
    # the removed residuals never come back: another optimisation leaves them out of the sums
    n_active_before = int(keep.sum())
    ow.linearize_all(update_th=False); ow.apply_res()
    assert ow.accumulate(0)["resInA"] <= n_active_before
    E2, removed2 = ow.finish_optimize()
    assert not np.intersect1d(removed, removed2).size


def test_oob_keeps_old_energy(orc, synth):
    W = synth.make_window(nf=2, npts=40, seed=17)
    W["res_state"] = np.ones(len(W["res_point"]), np.int32)  # all OOB
    W["res_energy"] = np.full(len(W["res_point"]), 3.5, np.float32)
    ow = orc.Window(W)
    E = ow.linearize_all(update_th=False)
    assert abs(E - 3.5 * ow.nres) < 1e-6
    assert np.all(ow.res_outputs(False)["newState"] == 1)


def test_pyramid_numpy_vs_oracle(orc, synth):
    rng = np.random.default_rng(0)
    img = (rng.random((480, 640)) * 255).astype(np.float32)
    a = synth.make_pyramid(img, synth.pyr_levels(640, 480))
    b = orc.make_images(img, (320, 320, 319.5, 239.5))
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    assert orc.lib().orc_pyr_levels(512, 512, 0) == 4
