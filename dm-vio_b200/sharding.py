"""Point sharding of one sliding window over ranks.

Every accumuland of the hot path is a sum over points and a point's Schur contribution involves only its own residuals
(reference: OptimizationBackend/AccumulatedSCHessian.cpp:L62-76), so points — with all their residuals — are the shard unit.
Images and the per-pair tables are replicated; the stitched (8nf+4)^2 system is the only thing exchanged (one all-reduce).
"""
import numpy as np

POINT_KEYS = ("host", "u", "v", "idepth", "idepth_zero", "color", "weights", "hasDepthPrior")


def shard_points(host, rank, world):
    """Boolean mask of the points owned by `rank`: round-robin inside the host-ordered list, which keeps every shard ordered by
    host frame (required by dmv_ba_set_points) and balances both the point and the residual count per (host, rank)."""
    n = len(host)
    return (np.arange(n) % world) == rank


def shard_window(W, rank, world):
    """The sub-window of `rank`: its points and their residuals re-indexed; everything else shared."""
    if world == 1:
        return W
    keep = shard_points(W["host"], rank, world)
    newidx = -np.ones(len(keep), np.int64)
    newidx[keep] = np.arange(int(keep.sum()))
    S = dict(W)
    for k in POINT_KEYS:
        S[k] = W[k][keep]
    rk = keep[W["res_point"]]
    S["res_point"] = newidx[W["res_point"][rk]].astype(np.int32)
    S["res_target"] = W["res_target"][rk]
    for k in ("res_state", "res_energy"):
        if W.get(k) is not None:
            S[k] = W[k][rk]
    S["shard_res_index"] = np.nonzero(rk)[0]  # positions of this shard's residuals in the full residual list
    return S


def pack_system(HA, bA, Hsc, bsc, energy, counts):
    """One flat fp64 buffer per rank == the payload of the single all-reduce (2*(N^2+N) + 4 doubles)."""
    return np.concatenate([np.ravel(HA), np.ravel(bA), np.ravel(Hsc), np.ravel(bsc), [float(energy)], np.asarray(counts, np.float64)])


def unpack_system(buf, N):
    o, out = 0, {}
    for k, n in (("HA", N * N), ("bA", N), ("Hsc", N * N), ("bsc", N)):
        out[k] = buf[o:o + n].reshape((N, N) if n == N * N else (N,))
        o += n
    out["energy"] = float(buf[o])
    out["counts"] = buf[o + 1:]
    return out
