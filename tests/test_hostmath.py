"""Host-side tables (precalc, adjoints, priors) of the product's numpy host layer vs the oracle's independent C++ restatement."""
import numpy as np


def test_precalc_adjoints_match_oracle(orc, synth):
    import dmvio_b200.hostmath as hm
    W = synth.make_window(nf=5, npts=50, seed=3)
    ow = orc.Window(W)
    pc_o, pc_p = ow.precalc(), hm.precalc_table(W)
    np.testing.assert_allclose(pc_p, pc_o, rtol=0, atol=1e-12)  # same float32 operation order as the reference (sequential products, cofactor inverse)
    adH_o, adT_o = ow.adjoints()
    adH_p, adT_p = hm.adjoints(W)
    np.testing.assert_allclose(adH_p, adH_o, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(adT_p, adT_o, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(hm.calib8(W["K"]), ow.calib()["k8"], rtol=1e-7)
    # priors (accumulateLF_MT with no linearised residuals == priors only)
    ow.linearize_all(); ow.apply_res()
    a = ow.accumulate(1)
    HL, bL = hm.prior_system(W)
    np.testing.assert_allclose(HL, a["HL"], rtol=1e-12)
    np.testing.assert_allclose(bL, a["bL"], rtol=1e-9, atol=1e-6)
    # same dense solve
    x_o, _, _ = ow.solve(0, 1e-5, 1)
    x_p = hm.solve_reduced(a["HA"], a["bA"], a["Hsc"], a["bsc"], HL, bL, lam=1e-5)
    # LU vs LDLT on a 44x44 system with condition ~1e12 (50 points only): cond * eps(fp64) ~ 1e-4 is the attainable agreement
    assert np.linalg.norm(x_p - x_o) <= 1e-3 * np.linalg.norm(x_o)
