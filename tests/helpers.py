"""Shared helpers for the parity tests: load one synthetic window into the CPU oracle and into the CUDA product."""
import numpy as np


def calib8_from_oracle(ow):
    return ow.calib()["k8"]


def product_ba_from_oracle(capi, W, ow, chunk_points=0, device=0, max_points=None):
    """Feeds the C ABI with the host-side tables computed by the oracle (precalc, adjoints, TH)."""
    ba = capi.BA(W["w"], W["h"], max_frames=max(2, W["nf"]), max_points=max_points or len(W["host"]), device=device, chunk_points=chunk_points)
    for k in range(W["nf"]):
        ba.upload_frame(k, W["dI"][k])
    ba.set_window(W["nf"])
    ba.set_points(W["host"], W["u"], W["v"], W["idepth"], W["idepth_zero"], W["color"], W["weights"])
    ba.set_residuals(W["res_point"], W["res_target"], W.get("res_state"), W.get("res_energy"))
    adH, adT = ow.adjoints()
    ba.set_adjoints(adH, adT)
    ba.set_state(calib8_from_oracle(ow), ow.precalc(), ow.frame_tables()["frameEnergyTH"])
    return ba


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def activation_case(synth, orc, seed=9, nf=5, n=2000, **kw):
    """A window plus immature points with depth intervals (half of them wide, half narrow) for the point-activation tests."""
    W = synth.make_window(nf=nf, npts=50, seed=seed, trans=0.06, rot=0.01, **kw)
    w, h = W["w"], W["h"]
    rng = np.random.default_rng(seed)
    host = np.sort(rng.integers(0, nf, n)).astype(np.int32)
    u, v = rng.integers(10, w - 10, n), rng.integers(10, h - 10, n)
    parts = [orc.ip_init(W["dI"][hh], w, h, u[host == hh], v[host == hh]) for hh in range(nf)]
    P = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    idt = (0.5 * (1 + 0.02 * rng.standard_normal(n))).astype(np.float32)
    wide = rng.random(n) < 0.5
    P["idepth_min"] = (idt * np.where(wide, 0.7, 0.97)).astype(np.float32)
    P["idepth_max"] = (idt * np.where(wide, 1.4, 1.03)).astype(np.float32)
    bad = rng.random(n) < 0.03            # a few hopeless intervals: far from the truth
    P["idepth_min"][bad] *= 3; P["idepth_max"][bad] *= 3
    return W, host, P


def init_points(rng, w, h, levels, counts):
    """points per level the way CoarseInitializer::setFirst lays them out (integer pixel + 0.1 inside the pattern margin, row-major order), with
    brute-force 10 nearest neighbours and the nearest parent one level up (CoarseInitializer::makeNN's outputs, without its kd-tree)"""
    pts = []
    for l in range(levels):
        wl, hl = w >> l, h >> l
        n = counts[l]
        x = rng.integers(3, wl - 4, 4 * n); y = rng.integers(3, hl - 4, 4 * n)
        xy = np.unique(np.stack([y, x], 1), axis=0)                       # row-major order, no duplicates
        xy = xy[np.sort(rng.choice(len(xy), min(n, len(xy)), replace=False))]
        pts.append(dict(u=(xy[:, 1] + 0.1).astype(np.float32), v=(xy[:, 0] + 0.1).astype(np.float32), type=np.ones(len(xy), np.float32)))
    for l in range(levels):
        p = pts[l]
        P = np.stack([p["u"], p["v"]], 1).astype(np.float64)
        d = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d, np.inf)
        k = min(10, len(P) - 1)
        nb = np.full((len(P), 10), -1, np.int32)
        nb[:, :k] = np.argsort(d, axis=1, kind="stable")[:, :k]
        p["neighbours"] = nb
        if l + 1 < levels:
            Q = np.stack([pts[l + 1]["u"], pts[l + 1]["v"]], 1).astype(np.float64)
            dq = ((0.5 * P[:, None, :] - Q[None, :, :]) ** 2).sum(-1)
            p["parent"] = np.argmin(dq, axis=1).astype(np.int32)
        else:
            p["parent"] = np.full(len(P), -1, np.int32)
    return pts
