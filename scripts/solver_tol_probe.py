#!/usr/bin/env python
"""Device Newton solver: evaluations per iteration and the cost in accuracy of the tolerance below which a step is
taken without another evaluation (fit.solver_tol), on fresh config-3 fits.  python scripts/solver_tol_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdiff_b200 as fcdiff

(N, H, U) = (400, 500, 500)
(_, _, _, _, b, bt) = fcdiff.UnsharedRegionModel().sample_device(N, H, U)


def run(tol, iters=10):
    f = fcdiff.fit.UnsharedRegionFit(); f.model = fcdiff.UnsharedRegionModel(); f.model.eta += 0.1
    (f.b, f.bt) = (b, bt); f.max_iters = iters; f.rel_tol = -1.0; f.solver_tol = tol
    torch.cuda.synchronize(); t0 = time.perf_counter(); f.run(); torch.cuda.synchronize()
    return f, (time.perf_counter() - t0) * 1e3

ref, _ = run(1e-10)
for tol in (1e-7, 1e-6, 1e-5, 1e-4):
    run(tol)
    (f, ms) = run(tol)
    th = np.array([f.model.pi, f.model.eta, f.model.epsilon]); th0 = np.array([ref.model.pi, ref.model.eta, ref.model.epsilon])
    print("tol %.0e: %.3f ms/iter  evals %s  rel theta %.2e  rel energy %.2e  max |dlqF| %.2e |dlqR| %.2e" % (
        tol, ms / 10, f.n_objective_evals, np.max(np.abs(th - th0) / th0),
        np.max(np.abs((np.array(f.energy) - np.array(ref.energy)) / np.array(ref.energy))),
        np.max(np.abs(np.exp(f._lq_F) - np.exp(ref._lq_F))), np.max(np.abs(np.exp(f._lq_R) - np.exp(ref._lq_R)))))
