"""Host-side (CPU, tiny) quantities that the reference also computes on the host and that the C ABI takes as inputs:
FrameFramePrecalc::set (src/dso/FullSystem/HessianBlocks.cpp:L193-223), EnergyFunctional::setAdjointsF
(src/dso/OptimizationBackend/EnergyFunctional.cpp:L48-108) and the Jacobi-preconditioned dense solve of
solveSystemF's no-GTSAM branch (EnergyFunctional.cpp:L909-916, L971-973).  numpy only; used by bench.py and the
Python-level tests.  (The C++ mirror of the same interface lives in host/.)"""
import numpy as np

from .synth import SCALE_A, SCALE_B, SCALE_C, SCALE_F, hat, se3_exp, se3_inv, se3_mul


def calib8(K_scaled):
    """CalibHessian::value_scaledf / value_scaledi (HessianBlocks.h:L356-371)."""
    k = np.asarray(K_scaled, np.float64).astype(np.float32)
    return np.array([k[0], k[1], k[2], k[3], np.float32(1) / k[0], np.float32(1) / k[1], -k[2] / k[0], -k[3] / k[1]], np.float32)


def frame_poses(W, state=None):
    """PRE_worldToCam = exp(state_scaled[:6]) * worldToCam_evalPT (HessianBlocks.h:L172-186)."""
    state = W["state"] if state is None else state
    out = []
    for k in range(W["nf"]):
        Re, te = se3_exp(state[k, :6])  # SCALE_XI_* = 1
        out.append(se3_mul(Re, te, W["R_eval"][k], W["t_eval"][k]))
    return out


def aff_from_to(expF, expT, aF, bF, aT, bT):
    """AffLight::fromToVecExposure (util/NumType.h:L174-186)."""
    if expF == 0 or expT == 0:
        expF = expT = 1.0
    a = np.exp(aT - aF) * expT / expF
    return a, bT - a * bF


def inv3_cofactor_f32(M):
    """3x3 float32 inverse the way Eigen evaluates it for fixed sizes: cofactors times one reciprocal of the determinant
    (FrameFramePrecalc::set computes K.inverse() like this, HessianBlocks.cpp:L217); every step rounded to float32."""
    M = np.asarray(M, np.float32)
    f = np.float32

    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return f(f(M[i1, j1] * M[i2, j2]) - f(M[i1, j2] * M[i2, j1]))

    c0, c1, c2 = cof(0, 0), cof(1, 0), cof(2, 0)
    det = f(f(f(c0 * M[0, 0]) + f(c1 * M[1, 0])) + f(c2 * M[2, 0]))
    invdet = f(f(1.0) / det)
    out = np.zeros((3, 3), np.float32)
    out[0, 0], out[0, 1], out[0, 2] = f(c0 * invdet), f(c1 * invdet), f(c2 * invdet)
    out[1, 0], out[1, 1], out[1, 2] = f(cof(0, 1) * invdet), f(cof(1, 1) * invdet), f(cof(2, 1) * invdet)
    out[2, 0], out[2, 1], out[2, 2] = f(cof(0, 2) * invdet), f(cof(1, 2) * invdet), f(cof(2, 2) * invdet)
    return out


def mm3_f32(A, B):
    """3x3 float32 product with the sequential inner-product order of the reference's fixed-size Eigen product, ((a0*b0 + a1*b1) + a2*b2),
    every step rounded to float32 (numpy's matmul may reorder / fuse)."""
    A = np.asarray(A, np.float32); B = np.asarray(B, np.float32)
    return ((A[:, 0:1] * B[0:1, :] + A[:, 1:2] * B[1:2, :]) + A[:, 2:3] * B[2:3, :]).astype(np.float32)


def precalc_table(W, state=None, K_scaled=None):
    """nf*nf*32 float32, index h*nf+t: KRKi[9] Kt[3] R0[9] t0[3] aff[2] b0 pad[5]."""
    nf = W["nf"]
    state = W["state"] if state is None else state
    K_scaled = W["K"] if K_scaled is None else K_scaled
    k8 = calib8(K_scaled)
    K = np.zeros((3, 3), np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = k8[0], k8[1], k8[2], k8[3], 1
    Ki = inv3_cofactor_f32(K)
    cur = frame_poses(W, state)
    out = np.zeros((nf * nf, 32), np.float32)
    for h in range(nf):
        Rh0i, th0i = se3_inv(W["R_eval"][h], W["t_eval"][h])
        Rhi, thi = se3_inv(*cur[h])
        for t in range(nf):
            R0, t0 = se3_mul(W["R_eval"][t], W["t_eval"][t], Rh0i, th0i)
            R, tt = se3_mul(cur[t][0], cur[t][1], Rhi, thi)
            Rf = R.astype(np.float32)
            q = out[h * nf + t]
            q[0:9] = mm3_f32(mm3_f32(K, Rf), Ki).reshape(-1)
            q[9:12] = mm3_f32(K, tt.astype(np.float32).reshape(3, 1)).reshape(-1)
            q[12:21] = R0.astype(np.float32).reshape(-1)
            q[21:24] = t0.astype(np.float32)
            a, b = aff_from_to(W["exposure"][h], W["exposure"][t], state[h, 6] * SCALE_A, state[h, 7] * SCALE_B, state[t, 6] * SCALE_A,
                               state[t, 7] * SCALE_B)
            q[24], q[25] = np.float32(a), np.float32(b)
            q[26] = np.float32(W["state_zero"][h, 7] * SCALE_B)
    return out


def adjoints(W):
    """adHost/adTarget, (nf*nf, 8, 8) float64, index h + t*nf."""
    nf = W["nf"]
    adH = np.zeros((nf * nf, 8, 8))
    adT = np.zeros((nf * nf, 8, 8))
    sc = np.array([1, 1, 1, 1, 1, 1, SCALE_A, SCALE_B], np.float64)
    for h in range(nf):
        Rhi, thi = se3_inv(W["R_eval"][h], W["t_eval"][h])
        for t in range(nf):
            R, tt = se3_mul(W["R_eval"][t], W["t_eval"][t], Rhi, thi)
            Adj = np.zeros((6, 6))
            Adj[:3, :3] = R
            Adj[3:, 3:] = R
            Adj[:3, 3:] = hat(tt) @ R
            AH, AT = np.eye(8), np.eye(8)
            AH[:6, :6] = -Adj.T
            a0, _ = aff_from_to(W["exposure"][h], W["exposure"][t], W["state_zero"][h, 6] * SCALE_A, W["state_zero"][h, 7] * SCALE_B,
                                W["state_zero"][t, 6] * SCALE_A, W["state_zero"][t, 7] * SCALE_B)
            a0 = float(np.float32(a0))
            AT[6, 6], AH[6, 6], AT[7, 7], AH[7, 7] = -a0, a0, -1.0, a0
            adH[h + t * nf] = AH * sc[:, None]
            adT[h + t * nf] = AT * sc[:, None]
    return adH, adT


def solve_reduced(HA, bA, Hsc, bsc, HL=None, bL=None, HM=None, bM_top=None, lam=1e-5):
    """EnergyFunctional::solveSystemF, default solver mode without GTSAM (EnergyFunctional.cpp:L909-916, L971-973)."""
    N = HA.shape[0]
    H = HA.copy()
    b = bA - bsc
    for M, v in ((HL, bL), (HM, bM_top)):
        if M is not None:
            H = H + M
            b = b + v
    H[np.diag_indices(N)] *= (1 + lam)
    H = H - Hsc / (1 + lam)
    s = 1.0 / np.sqrt(np.diag(H) + 10)
    return s * np.linalg.solve(s[:, None] * H * s[None, :], s * b)


def prior_system(W, state=None, cPrior=5e9, initialTransPrior=1e10, initialRotPrior=1e11, initialAffA=1e14, initialAffB=1e14, affA=1e12, affB=1e8):
    """accumulateLF_MT's prior part (AccumulatedTopHessian.cpp:L292-302) with FrameHessian::getPrior (HessianBlocks.h:L262-298)."""
    nf = W["nf"]
    N = 8 * nf + 4
    state = W["state"] if state is None else state
    # the reference's settings are C floats (util/settings.cpp:L67-73)
    cPrior, initialTransPrior, initialRotPrior, initialAffA, initialAffB, affA, affB = [
        float(np.float32(v)) for v in (cPrior, initialTransPrior, initialRotPrior, initialAffA, initialAffB, affA, affB)]
    HL = np.zeros((N, N))
    bL = np.zeros(N)
    HL[np.arange(4), np.arange(4)] = cPrior  # cDeltaF = 0 for an unmoved calibration
    for f in range(nf):
        p = np.zeros(8)
        if W["frameID"][f] == 0:
            p[:3], p[3:6], p[6], p[7] = initialTransPrior, initialRotPrior, initialAffA, initialAffB
        else:
            p[6], p[7] = affA, affB
        idx = 4 + 8 * f + np.arange(8)
        HL[idx, idx] += p
        bL[idx] += p * state[f, :8]
    return HL, bL


def trace_tables(W, host, new, state=None):
    """Operands of ImmaturePoint::traceOn as FullSystem::traceNewCoarse builds them (FullSystem.cpp:L548-561), in the reference's
    float operation order: KRKi = K * R.cast<float>() * K.inverse(), Kt = K * t.cast<float>(), aff = fromToVecExposure(...).cast<float>()."""
    state = W["state"] if state is None else state
    k8 = calib8(W["K"])
    K = np.zeros((3, 3), np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = k8[0], k8[1], k8[2], k8[3], 1
    Ki = inv3_cofactor_f32(K)
    cur = frame_poses(W, state)
    Rhi, thi = se3_inv(*cur[host])
    R, t = se3_mul(cur[new][0], cur[new][1], Rhi, thi)
    KRKi = mm3_f32(mm3_f32(K, R.astype(np.float32)), Ki)
    Kt = mm3_f32(K, t.astype(np.float32).reshape(3, 1)).reshape(-1)
    a, b = aff_from_to(W["exposure"][host], W["exposure"][new], state[host, 6] * SCALE_A, state[host, 7] * SCALE_B, state[new, 6] * SCALE_A,
                       state[new, 7] * SCALE_B)
    return KRKi, Kt, np.array([a, b], np.float32)
