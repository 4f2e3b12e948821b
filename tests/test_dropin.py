"""Drop-in proof: the REFERENCE's own translation units, compiled from /root/reference, linked with oracle/dropin_stubs.cpp so
that its hot-path members (CoarseTracker::calcRes / calcGSSSE, EnergyFunctional::accumulateAF_MT / accumulateSCF_MT / resubstituteF_MT,
CoarseInitializer::calcResAndGS) forward to
the C ABI of include/dmvio_b200.h (oracle/ref_build.sh dropin -> oracle/_ref/libdso_ref_dropin.so).  Everything around those members — the object
graph, PointFrameResidual::linearize, EnergyFunctional::solveSystemF with its dense solve, CoarseTracker::trackNewestCoarse with its LM loop — is
the reference's unmodified code.  The same harness calls run against the unmodified oracle/_ref/libdso_ref.so in this process and against the
drop-in build in a child process (the reference keeps process-global state); the results must agree within the CUDA path's tolerances.

  -m "not gpu": the drop-in build runs on the CPU stand-in of the C ABI (oracle/mock_capi.cpp): checks the stubs' plumbing.
  -m gpu:       the drop-in build runs on dm-vio_b200/libdmvio_b200.so: the reference's code driving the CUDA kernels."""
import os
import subprocess
import sys

import numpy as np
import pytest

import dropin_worker as dw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libdso_ref_dropin.so")


def _ensure_dropin():
    if not os.path.exists(DROPIN):
        if not os.path.isdir("/root/reference/src/dso"):
            pytest.skip("oracle/_ref/libdso_ref_dropin.so not built and /root/reference absent")
        if not os.path.exists(os.path.join(ROOT, "dm-vio_b200", "libdmvio_b200.so")):
            pytest.skip("dm-vio_b200/libdmvio_b200.so not built (run __graft_entry__.build())")
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh"), "dropin"])


def _run_dropin(case, tmp_path, mock):
    _ensure_dropin()
    env = dict(os.environ)
    env["DMV_REF_LIB"] = DROPIN
    if mock:  # the C ABI served by the CPU stand-in: a directory in which libdmvio_b200.so IS oracle/libhost_on_oracle.so, ahead of the RUNPATH
        d = tmp_path / "mocklib"
        d.mkdir(exist_ok=True)
        link = d / "libdmvio_b200.so"
        if not link.exists():
            os.symlink(os.path.join(ROOT, "oracle", "libhost_on_oracle.so"), link)
        env["LD_LIBRARY_PATH"] = str(d) + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    out = tmp_path / (case.replace(":", "_") + ".npz")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_worker.py"), case, str(out)], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    return dict(np.load(out))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _check_ba(d, r):
    # both sides ran the reference's own PointFrameResidual::linearize on the host objects: identical
    assert d["E"] == r["E"] and d["resInA"] == r["resInA"]
    for k in ("acc_HL", "acc_bL"):   # accumulateLF_MT is not replaced
        np.testing.assert_array_equal(d[k], r[k])
    # the replaced members: fp64-accumulating CUDA path vs the reference's fp32 SSE accumulators (same bounds as tests/test_ref_pin.py's
    # oracle-precision-1 check and tests/test_gpu_ba.py)
    for k, tol in (("acc_HA", 3e-6), ("acc_Hsc", 3e-6), ("acc_bA", 1e-4), ("acc_bsc", 1e-4)):
        assert _rel(d[k], r[k]) < tol, (k, _rel(d[k], r[k]))
    for k in ("pt_Hdd", "pt_bd", "pt_Hcd", "pt_HdiF", "pt_bdSumF"):   # per-point fp32 sums: the bounds of tests/test_gpu_ba.py
        np.testing.assert_allclose(d[k], r[k], rtol=2e-3, atol=2e-4 * np.abs(r[k]).max(), err_msg=k)
        assert _rel(d[k], r[k]) < 2e-4, (k, _rel(d[k], r[k]))
    # solveSystemF (the reference's code on both sides) on top of the replaced accumulators, then the replaced resubstituteF_MT.
    # The reduced system is ill-conditioned (gauge directions): the solution amplifies the accumulators' 1e-7 differences
    assert _rel(d["HS"], r["HS"]) < 3e-6 and _rel(d["bS"], r["bS"]) < 1e-4
    assert _rel(d["x"], r["x"]) < 5e-3, _rel(d["x"], r["x"])
    assert _rel(d["step"], r["step"]) < 1e-2, _rel(d["step"], r["step"])
    assert abs(d["E_hot"] - r["E_hot"]) <= 3e-4 * abs(r["E_hot"])


def _check_ct(d, r):
    L = int(r["levels"])
    assert int(d["levels"]) == L
    for l in range(L):
        for ci in range(2):
            np.testing.assert_allclose(d[f"res_{l}_{ci}"], r[f"res_{l}_{ci}"], rtol=2e-5, atol=1e-6)   # calcRes Vec6 (counts equal, energies fp32 sums)
            # the bounds of tests/test_gpu_coarse.py: H 2e-5, b 2e-4 (b is a sum of signed terms)
            np.testing.assert_allclose(d[f"H_{l}_{ci}"], r[f"H_{l}_{ci}"], rtol=0, atol=2e-5 * np.nanmax(np.abs(r[f"H_{l}_{ci}"]), initial=0) + 1e-30, equal_nan=True)
            np.testing.assert_allclose(d[f"b_{l}_{ci}"], r[f"b_{l}_{ci}"], rtol=0, atol=2e-4 * np.nanmax(np.abs(r[f"b_{l}_{ci}"]), initial=0) + 1e-30, equal_nan=True)
    # the reference's own trackNewestCoarse loop, every evaluation served by the stubs
    assert int(d["good"]) == int(r["good"])
    np.testing.assert_allclose(d["R"], r["R"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(d["t"], r["t"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(d["ab"], r["ab"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(d["lastResiduals"], r["lastResiduals"], rtol=2e-4, equal_nan=True)
    np.testing.assert_allclose(d["flow"], r["flow"], rtol=2e-3, atol=1e-4)


def _check_ci(d, r):
    L = int(r["levels"])
    for lvl in range(L):
        same = d[f"calc_{lvl}_good"] == r[f"calc_{lvl}_good"]
        assert same.mean() > 0.995
        np.testing.assert_allclose(d[f"calc_{lvl}_energy"][same], r[f"calc_{lvl}_energy"][same], rtol=3e-4, atol=1e-5)   # squares of differences of O(100) intensities
        if same.all():
            for k, tol in (("H", 5e-5), ("Hsc", 5e-5), ("b", 5e-4), ("bsc", 5e-4)):
                assert _rel(d[f"calc_{lvl}_{k}"], r[f"calc_{lvl}_{k}"]) < tol, (lvl, k, _rel(d[f"calc_{lvl}_{k}"], r[f"calc_{lvl}_{k}"]))
            np.testing.assert_allclose(d[f"calc_{lvl}_res"], r[f"calc_{lvl}_res"], rtol=5e-5)
    # the reference's own trackFrame on top: same decisions (ok / snapped / snappedAt / frameID), same poses to float accumulation order
    for k in range(6):
        np.testing.assert_array_equal(d[f"trk_{k}_state"], r[f"trk_{k}_state"], err_msg=f"frame {k}")
        np.testing.assert_allclose(d[f"trk_{k}_R"], r[f"trk_{k}_R"], rtol=0, atol=5e-5)
        np.testing.assert_allclose(d[f"trk_{k}_t"], r[f"trk_{k}_t"], rtol=0, atol=5e-4 * max(1e-2, np.abs(r[f"trk_{k}_t"]).max()))
        np.testing.assert_allclose(d[f"trk_{k}_ab"], r[f"trk_{k}_ab"], rtol=0, atol=2e-3)
    assert r["trk_5_state"][1] == 1   # the initialiser snapped
    same = d["final_isGood"] == r["final_isGood"]
    assert same.mean() > 0.99
    assert np.median(np.abs(d["final_idepth"][same] - r["final_idepth"][same]) / np.abs(r["final_idepth"][same])) < 2e-4


_CHECK = {"ba": _check_ba, "ct": _check_ct, "ci": _check_ci}


@pytest.fixture(scope="module")
def ref_results():
    from oracle import orc as _orc
    _orc.build()
    from oracle import ref as _ref
    if not _ref.available():
        pytest.skip("reference build not available")
    import dmvio_b200.synth as _synth
    cache = {}

    def get(case):
        if case not in cache:
            cache[case] = dw.run_case(_ref, _synth, case)
        return cache[case]
    return get


@pytest.mark.parametrize("case", ["ba:c1", "ba:states", "ct:small", "ci:many"])
def test_dropin_on_cpu_stand_in(ref_results, tmp_path, case):
    d, r = _run_dropin(case, tmp_path, mock=True), ref_results(case)
    _CHECK[case.split(":")[0]](d, r)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ba:c1", "ba:c3", "ba:states", "ct:small", "ct:vga", "ci:many"])
def test_dropin_on_cuda_library(ref_results, tmp_path, case):
    d, r = _run_dropin(case, tmp_path, mock=False), ref_results(case)
    _CHECK[case.split(":")[0]](d, r)
