"""Where a sharded EM iteration's time goes (run under torchrun, one rank per GPU; also works with one process).

Every phase of the loop body is bracketed by a device synchronisation + barrier, so the figures are the cost of the
phase when nothing overlaps it (their sum exceeds the free-running step, printed last)."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff
from fcdiff_b200 import _dev
from fcdiff_b200 import dist as fdist
import bench

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
shards = fdist.init_from_env("nccl") if world > 1 else None
N = bench.regions_for(world)
C = N * (N - 1) // 2
(H, U) = (bench.H_SUBJ, bench.U_SUBJ)
(c0, Cl) = (0, C) if shards is None else shards.span(C)
(_, _, _, _, b_dev, bt_dev) = fcdiff.UnsharedRegionModel().sample_device(N, H, U, c0=c0, C=Cl)
fit = fcdiff.fit.UnsharedRegionFit(); fit.model = fcdiff.UnsharedRegionModel(); fit.model.eta += 0.1
fit.b, fit.bt = b_dev, bt_dev
if shards is not None:
    fit.shards = shards; fit.n_edges = C
fit._init_lps(N, H, U); fit._update_lps(); fit._eval_energy()
for _ in range(3):
    bench.em_step(fit)


def sync():
    torch.cuda.synchronize()
    if shards is not None:
        dist.barrier()
    torch.cuda.synchronize()
    return time.perf_counter()


phases = [("K2+gather", fit._update_lq_F), ("K2b+gather", fit._update_lq_R),
          ("pi_gamma+bucket", lambda: fit._update_pi_gamma(True, True, fit._objective_context)),
          ("theta_sub", fit._update_theta_sub), ("lps+energy", lambda: (fit._update_lps(), fit._eval_energy()))]
acc = {k: 0.0 for (k, _) in phases}
R = 5
for _ in range(R):
    for (k, f) in phases:
        t0 = sync(); f(); t1 = sync()
        acc[k] += (t1 - t0) * 1e3 / R
t0 = sync()
for _ in range(R):
    bench.em_step(fit)
t1 = sync()
if shards is not None:
    # the collectives on their own
    (lqF, qF) = fit._mF.get_dev(); (lqR, qR) = fit._mR.get_dev()
    x = torch.zeros(4, dtype=torch.float64, device="cuda")
    res = {}
    for (k, f) in (("allgather_edges", lambda: shards.allgather_edges(lqF, qF, C)),
                   ("allgather_patients", lambda: shards.allgather_patients(lqR, qR, N, U)),
                   ("allreduce4+item", lambda: (dist.all_reduce(x), x.cpu()))):
        f(); a = sync()
        for _ in range(10):
            f()
        res[k] = (sync() - a) * 1e2
if rank == 0:
    print("world %d  N %d  C %d  nfev %.1f" % (world, N, C, float(np.mean(fit.n_objective_evals[-R:]))))
    for (k, _) in phases:
        print("%-18s %7.3f ms" % (k, acc[k]))
    print("sum of phases      %7.3f ms   free-running step %7.3f ms" % (sum(acc.values()), (t1 - t0) * 1e3 / R))
    if shards is not None:
        for (k, v) in res.items():
            print("%-18s %7.3f ms" % (k, v))
if shards is not None:
    dist.destroy_process_group()
