"""Where a FRESH sharded fit's time goes, phase by phase (set-up, initial energy, then every EM iteration split into
its step functions), each phase bracketed by a device synchronisation + barrier.  Run under torchrun (one rank per GPU)
or as one process.    python scripts/dist_setup_profile.py [iters]"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff
from fcdiff_b200 import fit as F
from fcdiff_b200 import dist as fdist
import bench

ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
shards = fdist.init_from_env("nccl") if world > 1 else None
N = bench.regions_for(world)
C = N * (N - 1) // 2
(H, U) = (bench.H_SUBJ, bench.U_SUBJ)
(c0, Cl) = (0, C) if shards is None else shards.span(C)
(_, _, _, _, b_dev, bt_dev) = fcdiff.UnsharedRegionModel().sample_device(N, H, U, c0=c0, C=Cl)


def sync():
    torch.cuda.synchronize()
    if shards is not None:
        dist.barrier()
    torch.cuda.synchronize()
    return time.perf_counter()


def fresh(profile):
    fit = F.UnsharedRegionFit(); fit.model = fcdiff.UnsharedRegionModel(); fit.model.eta += 0.1
    fit.b, fit.bt = b_dev, bt_dev
    if shards is not None:
        fit.shards = shards; fit.n_edges = C
    out = []

    def ph(name, f):
        t0 = sync() if profile else 0.0
        f()
        if profile:
            out.append((name, (sync() - t0) * 1e3))

    ph("init_lps+update_lps (inputs, planes)", lambda: (fit._init_lps(N, H, U), fit._update_lps()))
    ph("ensure_cache", fit._ensure_cache)
    ph("ensure_patient_planes", fit._ensure_patient_planes)
    ph("initial energy", lambda: fit.energy.append(fit._eval_energy()))
    for i in range(1, ITERS + 1):
        ph("it%d lq_F" % i, fit._update_lq_F)
        ph("it%d lq_R" % i, fit._update_lq_R)
        fit._more_iters = i < ITERS
        ph("it%d theta" % i, fit._update_theta)
        fit._more_iters = False
        ph("it%d lps+energy" % i, lambda: (fit._update_lps(), fit.energy.append(fit._eval_energy())))
    return out


fresh(False); fresh(False)
t0 = sync(); fresh(False); t1 = sync()
out = fresh(True)
if rank == 0:
    print("world %d  N %d  C %d: free-running fresh fit of %d iterations %.3f ms" % (world, N, C, ITERS, (t1 - t0) * 1e3))
    for (k, v) in out:
        print("  %-40s %7.3f ms" % (k, v))
    print("  sum %.3f ms" % sum(v for (_, v) in out))
if shards is not None:
    dist.destroy_process_group()
