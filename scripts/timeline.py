"""GPU timeline of one fresh fit (config 3 by default) from the per-call CUDA-event brackets:
every bracketed C-ABI call with its start offset, duration and the idle / unbracketed time before it.
    python scripts/timeline.py [regions] [subjects] [iters] [solver]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff                       # noqa: E402
from fcdiff_b200 import _dev                       # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
SOLVER = sys.argv[4] if len(sys.argv) > 4 else "newton"
m = fcdiff.UnsharedRegionModel()
(_, _, _, _, b, bt) = m.sample_device(N, S, S)


def fit_once(profile):
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.model = fcdiff.UnsharedRegionModel()
    fit.model.eta += 0.1
    fit.b, fit.bt = b, bt
    fit.max_iters = ITERS
    fit.rel_tol = -1.0
    fit.theta_solver = SOLVER
    fit.profile = profile
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit.run()
    torch.cuda.synchronize()
    return fit, (time.perf_counter() - t0) * 1e3


fit_once(None)
(_, ms_plain) = fit_once(None)
timers = _dev.KernelTimers()
origin = torch.cuda.Event(enable_timing=True)
origin.record()
(fit, ms) = fit_once(timers)
rows = []
for (name, evs) in timers.events.items():
    for (s, e, _) in evs:
        rows.append((origin.elapsed_time(s), s.elapsed_time(e), name))
rows.sort()
print("fit of %d iterations: %.3f ms without brackets, %.3f ms with; evals %s" % (ITERS, ms_plain, ms, fit.n_objective_evals))
end_prev = 0.0
busy = 0.0
for (t0, dur, name) in rows:
    print("%9.3f  +%7.3f gap  %8.3f ms  %s" % (t0, t0 - end_prev, dur, name))
    end_prev = t0 + dur
    busy += dur
print("bracketed %.3f ms of %.3f" % (busy, end_prev))
