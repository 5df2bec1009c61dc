"""Host-side anatomy of one K3b objective evaluation at config 3 (where the time between two kernels goes)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff
from fcdiff_b200 import _dev, _opt
import bench

(N, H, U) = (400, 500, 500)
(_, _, _, _, b, bt) = fcdiff.UnsharedRegionModel().sample_device(N, H, U)
fit = fcdiff.fit.UnsharedRegionFit(); fit.model = fcdiff.UnsharedRegionModel(); fit.model.eta += 0.1
fit.b, fit.bt = b, bt
fit._init_lps(N, H, U); fit._update_lps(); fit._eval_energy()
for _ in range(4):
    bench.em_step(fit)
fit._update_lq_F(); fit._update_lq_R(); fit._update_pi_gamma(True, True, fit._objective_context)
ctx = fit._objective_context()
(th, fn, head, tail, res) = (ctx['th'], ctx['fn'], ctx['head'], ctx['tail'], ctx['res'])
stream = ctx['stream']
x = np.array([fit.model.eta, fit.model.epsilon])
R = 200


def timeit(f, sync_each=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(R):
        f()
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / R * 1e6


print("launch only (host call, async)     %7.1f us  (GPU-bound when queued: kernel time)" % timeit(lambda: fn(*head, 1, *tail)))
t0 = time.perf_counter()
for _ in range(R):
    fn(*head, 1, *tail)
t_host = (time.perf_counter() - t0) / R * 1e6
torch.cuda.synchronize()
print("launch call host time              %7.1f us" % t_host)
print("launch + cuda sync                 %7.1f us" % timeit(lambda: fn(*head, 1, *tail), True))
print("launch + publish + wait            %7.1f us" % timeit(lambda: (fn(*head, 1, *tail), res.read(stream))))
print("publish + wait alone               %7.1f us" % timeit(lambda: res.read(stream)))
print("fit._objective(x)                  %7.1f us" % timeit(lambda: fit._objective(x)))
ev = [0]


def fun(xx):
    ev[0] += 1
    return fit._objective(xx)


torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    _opt.minimize_lbfgsb(fun, x, [1e-5, 1e-5], [1 - 1e-5, 1 - 1e-5])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) * 1e6
print("minimize_lbfgsb: %d evals, %.1f us per eval" % (ev[0], dt / ev[0]))
ev[0] = 0
t0 = time.perf_counter()
for _ in range(20):
    _opt.minimize_lbfgsb(lambda xx: (ev.__setitem__(0, ev[0] + 1), (float(np.sum((xx - 0.3) ** 2)), 2 * (xx - 0.3)))[1], x,
                         [1e-5, 1e-5], [1 - 1e-5, 1 - 1e-5])
dt = (time.perf_counter() - t0) * 1e6
print("minimize_lbfgsb on a host quadratic: %d evals, %.1f us per eval" % (ev[0], dt / max(ev[0], 1)))
