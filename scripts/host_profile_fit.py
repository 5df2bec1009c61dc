"""cProfile of ONE fresh fit.run() (config 3, 3 iterations) after two warm-up fits: where the host spends
the first iteration (set-up, allocations, waits).  Also counts cudaMalloc calls per fit."""
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


def one(iters=3):
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.model = fcdiff.UnsharedRegionModel()
    fit.model.eta += 0.1
    fit.b, fit.bt = b, bt
    fit.max_iters = iters
    fit.rel_tol = -1.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit.run()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for i in range(3):
    a0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    ms = one()
    print("fit %d: %.3f ms, cudaMalloc calls %d, reserved %.2f GB" % (
        i, ms, torch.cuda.memory_stats().get("num_device_alloc", 0) - a0, torch.cuda.memory_reserved() / 1e9))
for iters in (1, 2, 3, 10):
    print("iters %d: %.3f ms" % (iters, min(one(iters) for _ in range(3))))
pr = cProfile.Profile()
pr.enable()
one()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(30)
st.sort_stats("cumulative").print_stats(45)
