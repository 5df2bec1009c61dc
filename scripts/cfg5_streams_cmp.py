import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import fcdiff_b200 as fcdiff
from fcdiff_b200 import sweep
(N, H, U) = (400, 500, 500)
model = fcdiff.UnsharedRegionModel()
(_, _, _, _, b, bt) = model.sample_device(N, H, U)
corr = torch.cat([b, bt], dim=1)
labels = np.r_[np.zeros(H, bool), np.ones(U, bool)]
start = fcdiff.UnsharedRegionModel(); start.eta += 0.1
for streams in (1, 2, 3, 1, 2):
    sweep.permutation_sweep(corr, labels, 5, model=start, gather=False, streams=streams)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = sweep.permutation_sweep(corr, labels, 199, model=start, gather=False, streams=streams)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("streams %d: 200 replicas in %.3f s = %.1f replicas/s (E0 %.6f)" % (streams, dt, 200 / dt, res[0]["energy"][-1]), flush=True)
