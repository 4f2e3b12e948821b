"""Seeded synthetic sliding windows / tracking pairs for parity tests and bench.py.

Pure numpy/scipy; deliberately independent of both the CUDA product path and of oracle/.
The scene is a textured plane seen by nf keyframes with small relative motion, so that photometric
residuals are small-but-nonzero at the generated state (most residuals IN, a few OOB/OUTLIER), like a
converged DSO window.  Image pyramids follow the rule of FrameHessian::makeImages
(reference src/dso/FullSystem/HessianBlocks.cpp:L128-191), point colours/weights the rule of the
ImmaturePoint constructor (src/dso/FullSystem/ImmaturePoint.cpp:L36-62).
"""
import numpy as np
from scipy import ndimage

PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.int32)
SCALE_A, SCALE_B, SCALE_F, SCALE_C = 10.0, 1000.0, 50.0, 50.0


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def se3_exp(xi):
    """Sophus convention: xi = (upsilon, omega); returns R (3,3), t (3,)."""
    xi = np.asarray(xi, dtype=np.float64)
    ups, om = xi[:3], xi[3:]
    th = np.linalg.norm(om)
    Om = hat(om)
    if th < 1e-10:
        R = np.eye(3) + Om + 0.5 * Om @ Om
        V = R
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th**2 * (Om @ Om)
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * Om + (th - np.sin(th)) / th**3 * (Om @ Om)
    return R, V @ ups


def se3_mul(Ra, ta, Rb, tb):
    return Ra @ Rb, Ra @ tb + ta


def se3_inv(R, t):
    return R.T, -R.T @ t


def pyr_levels(w, h, force=0):
    """util/globalCalib.cpp:L49-55"""
    if force:
        return force
    lv, wl, hl = 1, w, h
    while wl % 2 == 0 and hl % 2 == 0 and wl * hl > 5000 and lv < 6:
        wl //= 2
        hl //= 2
        lv += 1
    return lv


def level_K(K, lvl):
    fx, fy, cx, cy = K
    return (fx * 0.5**lvl, fy * 0.5**lvl, (cx + 0.5) / (1 << lvl) - 0.5, (cy + 0.5) / (1 << lvl) - 0.5)


def make_pyramid(img, levels):
    """img: (h,w) float32 -> list of (h_l, w_l, 3) float32 arrays [I, dx, dy]; border rows keep zero gradients."""
    out = []
    cur = np.ascontiguousarray(img, dtype=np.float32)
    for lvl in range(levels):
        if lvl > 0:
            p = out[-1][:, :, 0]
            hl, wl = p.shape[0] // 2, p.shape[1] // 2
            cur = (np.float32(0.25) * (p[0:2 * hl:2, 0:2 * wl:2] + p[0:2 * hl:2, 1:2 * wl:2] + p[1:2 * hl:2, 0:2 * wl:2] + p[1:2 * hl:2, 1:2 * wl:2])).astype(np.float32)
        hl, wl = cur.shape
        flat = cur.reshape(-1)
        d = np.zeros((hl * wl, 3), dtype=np.float32)
        d[:, 0] = flat
        idx = np.arange(wl, wl * (hl - 1))
        d[idx, 1] = np.float32(0.5) * (flat[idx + 1] - flat[idx - 1])
        d[idx, 2] = np.float32(0.5) * (flat[idx + wl] - flat[idx - wl])
        out.append(d.reshape(hl, wl, 3))
    return out


def interp33_bilin(dI, x, y):
    """getInterpolatedElement33BiLin (util/globalFuncs.h:L203-226) on channel 0, vectorised. dI: (h,w,3)."""
    ix = x.astype(np.int32)
    iy = y.astype(np.int32)
    tl = dI[iy, ix, 0]
    tr = dI[iy, ix + 1, 0]
    bl = dI[iy + 1, ix, 0]
    br = dI[iy + 1, ix + 1, 0]
    dx = (x - ix).astype(np.float32)
    dy = (y - iy).astype(np.float32)
    topInt = dx * tr + (1 - dx) * tl
    botInt = dx * br + (1 - dx) * bl
    leftInt = dy * bl + (1 - dy) * tl
    rightInt = dy * br + (1 - dy) * tr
    return dx * rightInt + (1 - dx) * leftInt, rightInt - leftInt, botInt - topInt


def make_texture(rng, size=1536, sigma=2.0):
    t = rng.random((size, size)).astype(np.float32)
    t = ndimage.gaussian_filter(t, sigma, mode="wrap")
    t += 0.5 * ndimage.gaussian_filter(rng.random((size, size)).astype(np.float32), 6.0, mode="wrap")
    t -= t.min()
    t *= 255.0 / t.max()
    return t.astype(np.float32)


def render_plane(tex, K, R, t, w, h, D, a=0.0, b=0.0, texel_per_unit=None):
    """Image of the world plane Z=D (textured with tex) seen by worldToCam (R,t); I = exp(a)*tex + b."""
    fx, fy, cx, cy = K
    if texel_per_unit is None:
        texel_per_unit = fx / D
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    rays = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], axis=-1)  # cam frame
    Rt = R.T
    rw = rays @ Rt.T  # R^T r
    cw = -Rt @ t  # camera centre in world
    lam = (D - cw[2]) / rw[..., 2]
    Xw = cw[None, None, :] + lam[..., None] * rw
    s = Xw[..., 0] * texel_per_unit + tex.shape[1] / 2
    tt = Xw[..., 1] * texel_per_unit + tex.shape[0] / 2
    img = ndimage.map_coordinates(tex, [tt, s], order=3, mode="wrap").astype(np.float32)
    depth = lam  # z in camera frame (ray z = 1)
    return (np.exp(a) * img + b).astype(np.float32), depth.astype(np.float32)


def make_window(nf=7, npts=2000, w=640, h=480, seed=1234, hosts="all_but_newest", idepth_noise=1e-2, state_noise=1e-3,
                trans=0.1, rot=0.03, D=2.0, random_images=False):
    """Returns a dict of numpy arrays describing one sliding window."""
    rng = np.random.default_rng(seed)
    K = (0.5 * w, 0.5 * w, 0.5 * w - 0.5, 0.5 * h - 0.5)
    tex = make_texture(rng)
    R_eval = np.zeros((nf, 3, 3))
    t_eval = np.zeros((nf, 3))
    state = np.zeros((nf, 10))
    state_zero = np.zeros((nf, 10))
    images, dI, depth = [], [], []
    for k in range(nf):
        xi = np.concatenate([rng.uniform(-trans, trans, 3), rng.uniform(-rot, rot, 3)])
        if k == 0:
            xi *= 0
        R_eval[k], t_eval[k] = se3_exp(xi)
        a_true, b_true = rng.uniform(-0.05, 0.05), rng.uniform(-5, 5)
        delta_true = rng.uniform(-1e-3, 1e-3, 6)
        # truth = exp(delta_true) * evalPT ; current state = truth + noise ; state_zero (FEJ) keeps pose part 0
        Rk, tk = se3_mul(*se3_exp(delta_true), R_eval[k], t_eval[k])
        if random_images:
            img = make_texture(rng, size=max(w, h), sigma=2.0)[:h, :w].copy()
            dep = np.full((h, w), D, dtype=np.float32)
        else:
            img, dep = render_plane(tex, K, Rk, tk, w, h, D, a_true, b_true)
        images.append(img)
        depth.append(dep)
        dI.append(make_pyramid(img, 1)[0].reshape(-1).copy())
        state[k, :6] = delta_true + rng.uniform(-state_noise, state_noise, 6)
        state[k, 6] = (a_true + rng.uniform(-2e-3, 2e-3)) / SCALE_A
        state[k, 7] = (b_true + rng.uniform(-0.2, 0.2)) / SCALE_B
        state_zero[k, 6] = state[k, 6] + rng.uniform(-1e-4, 1e-4)
        state_zero[k, 7] = state[k, 7] + rng.uniform(-1e-5, 1e-5)
    # points
    nhost = nf - 1 if hosts == "all_but_newest" else nf
    if hosts == "first":
        nhost = 1
    host = np.sort(rng.integers(0, nhost, npts)).astype(np.int32)
    u = rng.integers(8, w - 8, npts).astype(np.float32)
    v = rng.integers(8, h - 8, npts).astype(np.float32)
    idepth = np.zeros(npts, np.float32)
    color = np.zeros((npts, 8), np.float32)
    weights = np.zeros((npts, 8), np.float32)
    for k in range(nf):
        m = host == k
        if not m.any():
            continue
        dIk = dI[k].reshape(h, w, 3)
        idepth[m] = 1.0 / depth[k][v[m].astype(int), u[m].astype(int)]
        for j in range(8):
            c, gx, gy = interp33_bilin(dIk, u[m] + PATTERN[j, 0], v[m] + PATTERN[j, 1])
            color[m, j] = c
            weights[m, j] = np.sqrt(np.float32(2500.0) / (np.float32(2500.0) + gx * gx + gy * gy))
    idepth = (idepth * (1 + idepth_noise * rng.standard_normal(npts))).astype(np.float32)
    idepth_zero = (idepth * (1 + 1e-3 * rng.standard_normal(npts))).astype(np.float32)
    # residuals: every point -> every other frame (point-major, target ascending), like FullSystem.cpp:L1377-1390
    rp, rt = [], []
    for t in range(nf):
        pass
    pt_idx = np.repeat(np.arange(npts, dtype=np.int32), nf)
    tg_idx = np.tile(np.arange(nf, dtype=np.int32), npts)
    keep = tg_idx != host[pt_idx]
    res_point = pt_idx[keep].astype(np.int32)
    res_target = tg_idx[keep].astype(np.int32)
    return dict(
        w=w, h=h, nf=nf, K=np.array(K, np.float64), images=images, dI=dI, R_eval=R_eval, t_eval=t_eval, state=state, state_zero=state_zero,
        exposure=np.ones(nf, np.float32), frameEnergyTH=np.full(nf, 8 * 8 * 8, np.float32), frameID=np.arange(nf, dtype=np.int32),
        host=host, u=u, v=v, idepth=idepth, idepth_zero=idepth_zero, color=color, weights=weights,
        hasDepthPrior=np.zeros(npts, np.uint8), res_point=res_point, res_target=res_target, seed=seed,
    )


def make_tracking_pair(w=640, h=480, seed=4321, npts=2000, levels=0, D=2.0, trans=0.03, rot=0.01):
    """Reference keyframe + new frame for the coarse tracker (config 2). Returns pyramids, true relative pose,
    and per-point (Ku,Kv,new_idepth,HdiF) splat inputs for makeCoarseDepthL0."""
    rng = np.random.default_rng(seed)
    K = (0.5 * w, 0.5 * w, 0.5 * w - 0.5, 0.5 * h - 0.5)
    L = pyr_levels(w, h, levels)
    tex = make_texture(rng)
    Rr, tr = np.eye(3), np.zeros(3)
    xi = np.concatenate([rng.uniform(-trans, trans, 3), rng.uniform(-rot, rot, 3)])
    Rn, tn = se3_exp(xi)  # refToNew (ref = world)
    a_new, b_new = 0.03, 2.0
    img_r, dep_r = render_plane(tex, K, Rr, tr, w, h, D)
    img_n, _ = render_plane(tex, K, Rn, tn, w, h, D, a_new, b_new)
    pyr_r = make_pyramid(img_r, L)
    pyr_n = make_pyramid(img_n, L)
    u = rng.integers(6, w - 6, npts).astype(np.float32)
    v = rng.integers(6, h - 6, npts).astype(np.float32)
    nid = (1.0 / dep_r[v.astype(int), u.astype(int)] * (1 + 5e-3 * rng.standard_normal(npts))).astype(np.float32)
    HdiF = rng.uniform(1e-4, 1e-2, npts).astype(np.float32)
    return dict(w=w, h=h, K=np.array(K, np.float64), levels=L, pyr_ref=pyr_r, pyr_new=pyr_n, R_true=Rn, t_true=tn, a_new=a_new, b_new=b_new,
                Ku=u, Kv=v, new_idepth=nid, HdiF=HdiF, img_ref=img_r, img_new=img_n)
