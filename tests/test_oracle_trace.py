"""First-principles checks of the immature-point tracing oracle (ImmaturePoint::traceOn restated in oracle/orc_trace.cpp):
the depth interval of a well-traced point brackets the true inverse depth of the synthetic plane."""
import numpy as np


def test_traced_interval_brackets_true_depth(orc, synth):
    import dmvio_b200.hostmath as hm
    W = synth.make_window(nf=2, npts=10, seed=8, trans=0.06, rot=0.005, state_noise=0.0)
    w, h = W["w"], W["h"]
    rng = np.random.default_rng(0)
    n = 2000
    u, v = rng.integers(12, w - 12, n), rng.integers(12, h - 12, n)
    P = orc.ip_init(W["dI"][0], w, h, u, v)
    KRKi, Kt, aff = hm.trace_tables(W, 0, 1)
    Q = orc.ip_trace(P, W["dI"][1], w, h, KRKi, Kt, aff)
    good = Q["status"] == 0
    assert good.mean() > 0.7
    # true inverse depth of pixel (u,v) in the host frame: the synthetic scene is the plane Z = 2 seen by frame 0 (rendered per frame)
    S = synth.make_window(nf=2, npts=n, seed=8, trans=0.06, rot=0.005, state_noise=0.0, idepth_noise=0.0, hosts="first")
    # make_window samples its own random pixels; recompute the plane depth analytically for ours instead
    cur = hm.frame_poses(W, W["state"])
    R, t = cur[0]
    fx, fy, cx, cy = W["K"]
    rays = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones(n)], -1)
    rw = rays @ R            # R^T r
    cw = -R.T @ t
    lam = (2.0 - cw[2]) / rw[:, 2]
    id_true = 1.0 / lam
    inside = (Q["idepth_min"][good] <= id_true[good] * 1.02) & (Q["idepth_max"][good] >= id_true[good] * 0.98)
    assert inside.mean() > 0.9, inside.mean()
    width = (Q["idepth_max"][good] - Q["idepth_min"][good]) / id_true[good]
    assert np.median(width) < 0.2
