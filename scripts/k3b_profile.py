"""Wall-clock anatomy of the K3b code pass at config 3 (synchronised phases, steady state)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff
from fcdiff_b200 import _dev, _lib
import bench

(N, H, U) = (400, 500, 500)
(_, _, _, _, b, bt) = fcdiff.UnsharedRegionModel().sample_device(N, H, U)
fit = fcdiff.fit.UnsharedRegionFit(); fit.model = fcdiff.UnsharedRegionModel(); fit.model.eta += 0.1
fit.b, fit.bt = b, bt
fit._init_lps(N, H, U); fit._update_lps(); fit._eval_energy()
for _ in range(5):
    bench.em_step(fit)
lib = _lib.load()
inp = fit._ensure_cache()
(Cl, pitchU, c0) = (inp['Cl'], inp['pitchU'], inp['c0'])
(_, qF) = fit._mF.get_dev(); (_, qR) = fit._mR.get_dev()
(fstate, rstate) = (fit._mF.get_state(), fit._mR.get_state())
stream = _dev.stream()
tot = fit._result(2, tag="records"); res4 = fit._result(4, tag="elm")
planeStride = Cl * pitchU


def sync():
    torch.cuda.synchronize(); return time.perf_counter()


def code_plane():
    _lib.check(lib.fcd_code_plane(_dev.ptr(inp['P']), planeStride, Cl, U, pitchU, _dev.ptr(fstate[c0:]), _dev.ptr(rstate),
                                  rstate.shape[1], _dev.ptr(inp['nm']), _dev.ptr(inp['PsE']), _dev.ptr(inp['kcE']),
                                  _dev.ptr(inp['code']), inp['pitchQ'], _dev.ptr(inp['bk_counts']), _dev.ptr(inp['bk_offs']),
                                  _dev.ptr(tot.dev), stream))


def records(nd, nh):
    _lib.check(lib.fcd_code_records(_dev.ptr(inp['P']), planeStride, _dev.ptr(inp['PsE']), _dev.ptr(inp['code']), inp['pitchQ'],
                                    _dev.ptr(inp['L']), _dev.ptr(inp['Lsum']), Cl, U, pitchU, _dev.ptr(qF[c0 * 3:]),
                                    _dev.ptr(fstate[c0:]), _dev.ptr(qR), _dev.ptr(rstate), rstate.shape[1], N, _dev.ptr(inp['nm']),
                                    _dev.ptr(inp['bk_counts']), _dev.ptr(inp['bk_offs']), _dev.ptr(inp['bk_K']), _dev.ptr(inp['bk_KH']),
                                    _dev.ptr(inp['bk_rowoff']), _dev.ptr(inp['bk_D']), nd, _dev.ptr(inp['bk_H']), nh, _dev.ptr(res4.dev[3:]),
                                    _dev.ptr(_dev.workspace()), stream))


R = 20
code_plane(); (nd, nh) = (int(v) for v in tot.read(stream))
print("records nd = %d (%.2f %%), half records nh = %d (%.2f %% of the elements)" % (nd, 100.0 * nd / (Cl * U), nh, 100.0 * nh / (Cl * U)))
for (name, f) in (("code_plane (pstar + codes + scan)", code_plane), ("read nd", lambda: tot.read(stream)),
                  ("records (fill + weights)", lambda: records(nd, nh)), ("build_streams (all)", lambda: fit._build_streams(inp, res4))):
    f(); t0 = sync()
    for _ in range(R):
        f()
    print("%-36s %7.1f us" % (name, (sync() - t0) / R * 1e6))
