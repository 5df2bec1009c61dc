"""cProfile of consecutive EM iterations at config 3 (host-side overhead hunting)."""
import cProfile
import os
import pstats
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff

(N, H, U) = (400, 500, 500)
m = fcdiff.UnsharedRegionModel()
(_, _, _, _, b, bt) = m.sample_device(N, H, U)
fit = fcdiff.fit.UnsharedRegionFit()
fit.model = fcdiff.UnsharedRegionModel()
fit.model.eta += 0.1
fit.b, fit.bt = b, bt
fit._init_lps(N, H, U)
fit._update_lps()
fit._eval_energy()


def step():
    fit._update_lq_F()
    fit._update_lq_R()
    fit._update_theta()
    fit._update_lps()
    return fit._eval_energy()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print("ms/step %.3f  evals/step %.1f" % ((time.perf_counter() - t0) * 100, sum(fit.n_objective_evals[-10:]) / 10))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(25)
