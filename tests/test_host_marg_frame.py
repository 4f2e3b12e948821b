"""Host-side EnergyFunctional::marginalizeFrame of the product (dm-vio_b200/host/marg_frame.h, used by WindowBA::marginalizeFrame) against
the oracle's independent restatement, which tests/test_ref_pin.py pins to the reference's compiled code.  Pure host fp64: runs without a GPU."""
import numpy as np
import pytest

from helpers import rel


@pytest.mark.parametrize("nf,idx", [(7, 0), (7, 6), (4, 2), (2, 0), (8, 5)])
def test_marginalize_frame_hm_matches_oracle(orc, synth, nf, idx):
    import dmvio_b200.hostapi as hostapi
    W = dict(synth.make_window(nf=nf, npts=nf, seed=50 + nf))
    for k in ("host", "u", "v", "idepth", "idepth_zero", "color", "weights", "hasDepthPrior"):
        W[k] = W[k][:0]
    W["res_point"] = np.zeros(0, np.int32); W["res_target"] = np.zeros(0, np.int32)
    ow = orc.Window(W)
    ft = ow.frame_tables()
    rng = np.random.default_rng(7 * nf + idx)
    odim = 8 * nf + 4
    A = rng.standard_normal((odim, 2 * odim))
    scale = 10.0 ** rng.uniform(0, 6, odim)
    HM = (A @ A.T) * np.outer(scale, scale)
    bM = rng.standard_normal(odim) * scale
    H_g, b_g = hostapi.marginalize_frame_hm(HM, bM, nf, idx, ft["prior"][idx], ft["delta_prior"][idx])
    H_o, b_o = ow.marginalize_frame(idx, HM, bM)
    assert rel(H_g, H_o) < 1e-12 and rel(b_g, b_o) < 1e-12
    np.testing.assert_array_equal(H_g, H_g.T)
    # against first principles: marginalising a Gaussian = Schur complement of the permuted system with the frame prior added
    io = 4 + 8 * idx
    keep = np.r_[0:io, io + 8:odim]
    Hp = HM.copy(); bp = bM.copy()
    Hp[io:io + 8, io:io + 8] += np.diag(ft["prior"][idx]); bp[io:io + 8] += ft["prior"][idx] * ft["delta_prior"][idx]
    S = Hp[np.ix_(keep, keep)] - Hp[np.ix_(keep, np.arange(io, io + 8))] @ np.linalg.solve(Hp[io:io + 8, io:io + 8], Hp[np.ix_(np.arange(io, io + 8), keep)])
    sb = bp[keep] - Hp[np.ix_(keep, np.arange(io, io + 8))] @ np.linalg.solve(Hp[io:io + 8, io:io + 8], bp[io:io + 8])
    assert rel(H_g, S) < 1e-8 and rel(b_g, sb) < 1e-8


def test_grow_then_marginalize_roundtrip(orc, synth):
    """a frame that enters with zero rows/columns (EnergyFunctional::insertFrame) and leaves again without having been coupled to the others
    changes nothing: its own prior is marginalised away with it"""
    import dmvio_b200.hostapi as hostapi
    rng = np.random.default_rng(1)
    nf = 4
    odim = 8 * nf + 4
    A = rng.standard_normal((odim, odim + 5))
    HM = A @ A.T; bM = rng.standard_normal(odim)
    big = np.zeros((odim + 8, odim + 8)); big[:odim, :odim] = HM
    bbig = np.r_[bM, np.zeros(8)]
    H2, b2 = hostapi.marginalize_frame_hm(big, bbig, nf + 1, nf, np.full(8, 3.0), np.full(8, 0.5))
    assert rel(H2, HM) < 1e-12 and rel(b2, bM) < 1e-12


@pytest.mark.parametrize("cfg", [dict(nf=7, npts=50, seed=1234), dict(nf=2, npts=20, seed=3), dict(nf=8, npts=30, seed=99, rot=0.2, trans=0.5)],
                         ids=["nf7", "nf2", "nf8_big_motion"])
def test_host_nullspaces_and_orthogonalize_match_oracle(orc, synth, cfg):
    """host/nullspace.h (used by WindowBA::solveSystemF from iteration 2 on) vs the oracle, which tests/test_ref_pin.py pins to the reference's
    setStateZero / orthogonalize: nullspace vectors, projection of a random x, and its defining properties."""
    import dmvio_b200.hostapi as hostapi
    W = synth.make_window(**cfg)
    ow = orc.Window(W)
    rng = np.random.default_rng(cfg["seed"])
    x = rng.standard_normal(8 * W["nf"] + 4)
    ns_g, x_g = hostapi.nullspaces_orthogonalize(W["R_eval"], W["t_eval"], x)
    ns_o, x_o = ow.nullspaces(), ow.orthogonalize(x)
    np.testing.assert_allclose(ns_g, ns_o, rtol=0, atol=1e-9 * np.abs(ns_o).max())
    assert rel(x_g, x_o) < 1e-10
    Nn = ns_g / np.linalg.norm(ns_g, axis=1, keepdims=True)
    assert np.abs(Nn @ x_g).max() < 1e-9 * np.linalg.norm(x)
    # first principles: the projector onto the complement of span(N), from numpy's pseudo-inverse
    P = np.eye(len(x)) - Nn.T @ np.linalg.pinv(Nn.T)
    assert rel(x_g, P @ x) < 1e-9
    # the calibration and affine entries are untouched by pose / scale gauge directions
    idx = np.r_[0:4, [4 + 8 * f + k for f in range(W["nf"]) for k in (6, 7)]]
    np.testing.assert_allclose(x_g[idx], x[idx], rtol=0, atol=1e-12)
