"""CPU tests pinning the oracle's coarse-tracker restatement (CoarseTracker.cpp) by first principles."""
import numpy as np


def _setup(orc, synth, seed=4321, levels=0):
    T = synth.make_tracking_pair(seed=seed, levels=levels)
    ct = orc.CoarseTracker(T["w"], T["h"], T["K"], levels)
    n = ct.make_coarse_depth(T["Ku"], T["Kv"], T["new_idepth"], T["HdiF"], T["pyr_ref"])
    ct.set_new_frame(T["pyr_new"])
    return T, ct, n


def test_coarse_depth_lists(orc, synth):
    T, ct, n = _setup(orc, synth)
    assert ct.levels == 4
    counts = [len(ct.ref_points(l)["u"]) for l in range(ct.levels)]
    assert sum(counts) == n
    assert counts[0] > len(T["Ku"])          # dilation adds neighbours on level 0
    p0 = ct.ref_points(0)
    assert p0["idepth"].min() > 0
    # colours are the reference image sampled at integer pixels
    np.testing.assert_array_equal(p0["color"], T["pyr_ref"][0][p0["v"].astype(int), p0["u"].astype(int), 0])
    # 5 forced levels (BASELINE config 2)
    T5, ct5, _ = _setup(orc, synth, levels=5)
    assert ct5.levels == 5 and ct5.K(4)[1].tolist() == [40, 30]


def test_energy_minimal_at_true_pose(orc, synth):
    T, ct, _ = _setup(orc, synth)
    r_true = ct.calc_res(0, T["R_true"], T["t_true"], T["a_new"], T["b_new"])
    r_id = ct.calc_res(0, np.eye(3), np.zeros(3), 0.0, 0.0)
    assert r_true[0] / r_true[1] < 0.2 * r_id[0] / r_id[1]
    assert r_true[5] < 0.05


def test_gs_is_gradient_of_energy(orc, synth):
    """b = sum w J r / n is the gradient of 0.5*sum(huber energy)/n wrt a left pose increment, up to DSO's approximation
    (gradients come from the central-difference channel, not from the interpolant): direction within a few degrees,
    magnitude within the smoothing loss of the gradient channel."""
    T, ct, _ = _setup(orc, synth)
    R0, t0 = synth.se3_mul(*synth.se3_exp(np.array([0.004, -0.003, 0.002, 0.002, -0.001, 0.0015])), T["R_true"], T["t_true"])
    a, b = T["a_new"] + 0.01, T["b_new"] + 0.5
    for lvl, cos_min, ratio_min in ((0, 0.995, 0.85), (1, 0.98, 0.7)):
        ct.calc_res(lvl, R0, t0, a, b, cutoff=1e6)
        H, g = ct.calc_gs(lvl, a, b, 1)
        n = ct.warped().shape[1]
        assert np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-9 * np.abs(H).max()

        def energy(xi):
            Re, te = synth.se3_exp(xi)
            R1, t1 = synth.se3_mul(Re, te, R0, t0)
            return 0.5 * ct.calc_res(lvl, R1, t1, a, b, cutoff=1e6)[0]
        fd = np.zeros(6)
        for k in range(6):
            e = np.zeros(6); e[k] = 2e-4
            fd[k] = (energy(e) - energy(-e)) / 4e-4 / n
        cos = fd @ g[:6] / np.linalg.norm(fd) / np.linalg.norm(g[:6])
        ratio = np.linalg.norm(g[:6]) / np.linalg.norm(fd)
        assert cos > cos_min and ratio_min < ratio < 1.1, (lvl, cos, ratio)
    # affine part: d(0.5 E)/db = sum w r * (-1) ... in the scaled parametrisation b_out[7] = SCALE_B * sum(w*(-1)*r)/n
    lvl = 0
    ct.calc_res(lvl, R0, t0, a, b, cutoff=1e6)
    H, g = ct.calc_gs(lvl, a, b, 1)
    n = ct.warped().shape[1]
    eb = 0.5  # large step: calcRes sums the energy in fp32 (CoarseTracker.cpp:L363), the quadratic-ish b-dependence tolerates it
    fdb = (0.5 * ct.calc_res(lvl, R0, t0, a, b + eb, cutoff=1e6)[0] - 0.5 * ct.calc_res(lvl, R0, t0, a, b - eb, cutoff=1e6)[0]) / (2 * eb) / n
    assert abs(fdb * 1000.0 - g[7]) <= 0.02 * abs(g[7]) + 1e-3, (fdb * 1000.0, g[7])


def test_tracking_recovers_pose(orc, synth):
    T, ct, _ = _setup(orc, synth)
    res = ct.track(np.eye(3), np.zeros(3), 0.0, 0.0)
    assert res["good"]
    assert np.linalg.norm(res["t"] - T["t_true"]) < 0.1 * np.linalg.norm(T["t_true"]) + 2e-3
    assert np.abs(res["R"] - T["R_true"]).max() < 2e-3
    # a and b are strongly correlated (a*mean(I) + b): compare the brightness transfer at the mean intensity
    assert abs((np.exp(res["a"]) * 127 + res["b"]) - (np.exp(T["a_new"]) * 127 + T["b_new"])) < 0.5
    assert res["lastResiduals"][0] < 3.0
    # fp32-faithful accumulation takes the same path
    res32 = ct.track(np.eye(3), np.zeros(3), 0.0, 0.0, precision=0)
    assert np.abs(res32["R"] - res["R"]).max() < 1e-4
