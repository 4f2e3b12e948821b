"""In-kernel phase timeline of ba_point_kernel (DMV_DBG=16): globaltimer stamps per CTA."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DMV_DBG"] = os.environ.get("DMV_DBG", "16")
import numpy as np
import dmvio_b200.capi as capi, dmvio_b200.hostmath as hm, dmvio_b200.synth as synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W = synth.make_window(nf=7, npts=2000, seed=1234)
ba = capi.BA(640, 480, max_frames=7, max_points=2000, chunk_points=P)
for k in range(7): ba.upload_frame(k, W["dI"][k])
ba.set_window(7); ba.set_points(W["host"], W["u"], W["v"], W["idepth"], W["idepth_zero"], W["color"], W["weights"])
ba.set_residuals(W["res_point"], W["res_target"])
ba.set_adjoints(*hm.adjoints(W))
k8, pc, TH = hm.calib8(W["K"]), hm.precalc_table(W), W["frameEnergyTH"]
ba.set_state(k8, pc, TH); ba.linearize(); ba.apply_res()
acc = ba.accumulate(); HL, bL = hm.prior_system(W)
x = hm.solve_reduced(acc["HA"], acc["bA"], acc["Hsc"], acc["bsc"], HL, bL)
ba.backup_points()
L = capi.lib(); L.dmv_ba_debug_clocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
HAVE_X = int(os.environ.get('HAVE_X', '1'))
for flush in (True, False):
    ba.bench_device(x if HAVE_X else None, iters=5, flush_l2=flush)
    buf = np.zeros(16 * 600, np.uint64)
    n = L.dmv_ba_debug_clocks(ba.h, buf.ctypes.data_as(C.c_void_p), len(buf))
    t = buf[:n].reshape(-1, 16).astype(np.int64)
    ts = t[-8:]  # the nf+1 = 8 stitch CTAs
    t = t[:-8]
    t0 = t[:, 0].min()
    order = [0, 8, 9, 6, 3, 4, 5]
    names = {8: "chunk_known", 9: "adj_issued", 0: "start", 6: "prior_issued", 3: "phaseA+sync", 4: "phaseB", 5: "end"}
    print(f"P={P} flush={flush} have_x={HAVE_X} blocks={len(t)}  (ns relative to first CTA start; mean / max over CTAs)")
    prev = 0
    for k in order:
        print(f"  {names[k]:17s} mean {np.mean(t[:, k] - t0):9.0f}  max {np.max(t[:, k] - t0):9.0f}   dt_mean {np.mean(t[:, k] - t[:, prev]):8.0f}")
        prev = k
    print(f"  stitch (ns rel. to first point-CTA start, mean/max over {len(ts)} CTAs): resident {np.mean(ts[:,0]-t0):.0f}/{np.max(ts[:,0]-t0)}, dependency released {np.mean(ts[:,1]-t0):.0f}/{np.max(ts[:,1]-t0)}, "
          f"staged {np.mean(ts[:-1,2]-t0):.0f}/{np.max(ts[:-1,2]-t0)}, products done {np.mean(ts[:-1,3]-t0):.0f}/{np.max(ts[:-1,3]-t0)}, written {np.mean(ts[:-1,4]-t0):.0f}/{np.max(ts[:-1,4]-t0)}")
    print(f"  stitch CTA nf (calibration rows + tiles share): released {ts[-1,1]-t0}, done {ts[-1,4]-t0}")
