"""Multi-rank logic of the BA path on CPU (gloo, world_size 2): point sharding + the single all-reduce of the stitched system.

What is checked (the CUDA kernels are not involved; the per-shard systems come from the oracle, used here as the checker):
  * shards partition the points and residuals, each shard stays ordered by host frame;
  * sum over ranks of the per-shard [H_A b_A H_sc b_sc energy counters] == the unsharded window's (the shard-sum invariance that
    makes a plain ncclAllReduce(sum, double) the whole exchange step);
  * after the all-reduce every rank holds the identical system, hence solves to the identical x.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import rel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import dmvio_b200.synth as synth
    import dmvio_b200.hostmath as hm
    from dmvio_b200.sharding import shard_window, pack_system, unpack_system
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = synth.make_window(**cfg)
    S = shard_window(W, rank, world)
    ow = orc.Window(S)
    E = ow.linearize_all(update_th=False)
    st = ow.res_outputs(False)["newState"]
    ow.apply_res()
    a = ow.accumulate(1)
    buf = torch.from_numpy(pack_system(a["HA"], a["bA"], a["Hsc"], a["bsc"], E, [(st == 0).sum(), (st == 1).sum(), (st == 2).sum()]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    N = 8 * W["nf"] + 4
    tot = unpack_system(buf.numpy(), N)
    HL, bL = hm.prior_system(W)
    x = hm.solve_reduced(tot["HA"], tot["bA"], tot["Hsc"], tot["bsc"], HL, bL, lam=1e-5)
    q.put((rank, len(S["host"]), len(S["res_point"]), bool(np.all(np.diff(S["host"]) >= 0)), buf.numpy().copy(), x))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [dict(nf=4, npts=301, seed=5, w=160, h=120)], ids=["nf4_n301"])
def test_shard_allreduce_matches_unsharded(orc, synth, cfg):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W = synth.make_window(**cfg)
    assert sum(r[1] for r in res) == len(W["host"])
    assert sum(r[2] for r in res) == len(W["res_point"])
    assert all(r[3] for r in res)
    np.testing.assert_array_equal(res[0][4], res[1][4])  # identical on every rank after the all-reduce
    np.testing.assert_array_equal(res[0][5], res[1][5])
    ow = orc.Window(W)
    E = ow.linearize_all(update_th=False)
    st = ow.res_outputs(False)["newState"]
    ow.apply_res()
    a = ow.accumulate(1)
    from dmvio_b200.sharding import unpack_system
    tot = unpack_system(res[0][4], 8 * W["nf"] + 4)
    for k in ("HA", "bA", "Hsc", "bsc"):
        assert rel(tot[k], a[k]) < 1e-12, k
    assert abs(tot["energy"] - E) <= 1e-9 * abs(E)
    np.testing.assert_array_equal(tot["counts"], [(st == 0).sum(), (st == 1).sum(), (st == 2).sum()])


def _marg_worker(rank, world, port, cfg, flagged, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import dmvio_b200.synth as synth
    from dmvio_b200.sharding import shard_window, shard_points
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = synth.make_window(**cfg)
    mine = np.nonzero(shard_points(W["host"], rank, world))[0]          # global indices of this rank's points, in shard order
    local = np.nonzero(np.isin(mine, flagged))[0].astype(np.int32)       # the flagged ones, as shard-local indices
    S = shard_window(W, rank, world)
    o = orc.Window(S).marginalize(local, precision=1)
    N = 8 * W["nf"] + 4
    buf = torch.from_numpy(np.concatenate([o["H"].reshape(-1), o["b"], [float(o["resInM"])]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)                           # the exchange a sharded marginalisation needs: one sum of M - Msc | Mb - Mbsc
    q.put((rank, len(local), buf.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_marginalisation_sums_to_unsharded(orc, synth):
    """point marginalisation shards like the rest of the path: every rank marginalises the flagged points it owns, the sum of the per-rank
    M - Msc, Mb - Mbsc, resInM equals the unsharded EnergyFunctional::marginalizePointsF (per-point terms only: SURVEY.md section 8e)"""
    cfg = dict(nf=4, npts=301, seed=5, w=160, h=120)
    W = synth.make_window(**cfg)
    flagged = np.sort(np.random.default_rng(3).choice(len(W["host"]), 120, replace=False))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_marg_worker, args=(r, world, port, cfg, flagged, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[1] for r in res) == len(flagged)
    np.testing.assert_array_equal(res[0][2], res[1][2])
    N = 8 * W["nf"] + 4
    full = orc.Window(W).marginalize(flagged.astype(np.int32), precision=1)
    tot = res[0][2]
    assert rel(tot[:N * N].reshape(N, N), full["H"]) < 1e-12
    assert rel(tot[N * N:N * N + N], full["b"]) < 1e-11
    assert int(tot[-1]) == full["resInM"]


def test_shard_edge_cases(synth):
    from dmvio_b200.sharding import shard_window, shard_points
    W = synth.make_window(nf=3, npts=5, seed=2, w=96, h=64, hosts="all")
    # more ranks than points of some host; a rank may own zero points
    seen = np.zeros(5, int)
    for r in range(8):
        m = shard_points(W["host"], r, 8)
        seen += m
        S = shard_window(W, r, 8)
        assert len(S["res_point"]) == (len(S["host"]) * 2)
        if len(S["host"]):
            assert S["res_point"].max() < len(S["host"])
    np.testing.assert_array_equal(seen, 1)
    assert shard_window(W, 0, 1) is W


def _worker_adapter(rank, world, port, cfg, q):
    """one rank of the sharded C++ adapter on the CPU stand-in of the C ABI; the stand-in's system all-reduce and the adapter's allgather
    both go through gloo"""
    import ctypes as C
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import dmvio_b200.hostapi as hostapi
    import dmvio_b200.synth as synth
    from test_gpu_multi import _flow
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = C.CDLL(os.path.join(root, "oracle", "libhost_on_oracle.so"))
    hostapi._L = hostapi._bind(L)

    CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.c_void_p)

    def _allreduce(buf, n, user):
        a = np.ctypeslib.as_array(buf, shape=(n,))
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        a[:] = t.numpy()
    cb = CB(_allreduce)
    L.mock_set_allreduce.argtypes = [CB, C.c_void_p]
    L.mock_set_allreduce(cb, None)

    def allgather(data):
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return b"".join(o.numpy().tobytes() for o in outs)

    W = synth.make_window(**cfg)
    hw = hostapi.WindowBA(W, shard=dict(rank=rank, nranks=world, allgather=allgather, exchange="p2p"))
    out = _flow(hw, W["nf"])
    q.put((rank, out))
    dist.barrier()
    hw.close()
    dist.destroy_process_group()


def test_sharded_adapter_flow_matches_unsharded(orc, synth):
    """WindowBA::setSharding on the CPU: two ranks run the whole makeKeyFrame flow (optimize, tail with residual removal, flagPointsForRemoval,
    marginalizePointsF, marginalizeFrame, optimize again) on their shares; the stand-in of the C ABI sums the systems over the ranks where the
    CUDA library does it inside the launch.  Same decisions and state as the unsharded adapter; the ranks agree bit for bit."""
    import ctypes as C
    import subprocess
    import dmvio_b200.hostapi as hostapi
    from test_gpu_multi import _flow
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle"), "libhost_on_oracle.so"])
    world = 2
    cfg = dict(nf=5, npts=301, seed=31, state_noise=1e-3, hosts="all", w=160, h=120)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_adapter, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    saved = hostapi._L
    hostapi._L = hostapi._bind(C.CDLL(os.path.join(root, "oracle", "libhost_on_oracle.so")))
    try:
        ref = _flow(hostapi.WindowBA(synth.make_window(**cfg)), cfg["nf"])
    finally:
        hostapi._L = saved
    for rk in range(world):
        o = res[rk]
        assert (o["n1"], o["n2"], o["resInM"], o["nres"]) == (ref["n1"], ref["n2"], ref["resInM"], ref["nres"])
        np.testing.assert_array_equal(o["marg"], ref["marg"]); np.testing.assert_array_equal(o["drop"], ref["drop"])
        np.testing.assert_array_equal(o["removed"], ref["removed"])
        np.testing.assert_allclose(o["log1"], ref["log1"], rtol=1e-9)     # fp64 sums in another order only
        np.testing.assert_allclose(o["log2"], ref["log2"], rtol=1e-8)
        assert rel(o["HM"], ref["HM"]) < 1e-9 and rel(o["bM"], ref["bM"]) < 1e-8
        assert np.abs(o["st"] - ref["st"]).max() < 1e-9
        np.testing.assert_allclose(o["idepth"], ref["idepth"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(res[1]["st"], res[0]["st"])
    np.testing.assert_array_equal(res[1]["idepth"], res[0]["idepth"])
