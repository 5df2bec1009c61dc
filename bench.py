#!/usr/bin/env python
"""
bench.py -- edge-subject EM iterations per second of the variational-EM hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[2], the configuration the north_star target is
quoted on): Schaefer-400 atlas (79,800 edges) x 500 controls + 500 patients,
synthetic correlations drawn from the model defaults (fcdiff/model.py:33-38),
fit started from theta_true with eta + 0.1.  It fits one GPU.  With N > 1 ranks
the atlas grows so that every rank keeps ~79,800 edges (weak scaling; edges are
sharded, the small M-step / energy statistics are all-reduced, lq_F / lq_R are
all-gathered -- fcdiff_b200/dist.py).

A *step* is one pass of the reference's loop body (fcdiff/fit.py:76-80):
E-step q_F (K2), region weights + Gauss-Seidel sweep for q_R (K2b), pi/gamma
(K3a), the (eta, epsilon) solve: one code pass, then J objective evaluations with
the optimiser's step taken on the device (K3b), free energy (K4).

value  = C*(H+U)*K / t over FRESH FITS FROM THE UNIFORM START (fit.py:84-102) with
         inputs resident in HBM: the K timed steps are the iterations of
         ceil(K / 10) calls of ``fit.run()`` with ``max_iters = 10`` (the reference's
         default, fit.py:36) and the convergence test disabled (``rel_tol = -1``), so
         every fit runs all its iterations -- set-up (responsibility planes,
         dominant-state plane, initial energy) and the expensive first iterations,
         where no posterior is decided yet, are inside the timed region.  CUDA
         events, max over ranks.
steady_state = the same metric over consecutive iterations of ONE fit after the
         warm-up iterations (the regime of a long fit: ~90 % of the elements
         decided; round 1 reported this as `value`).
e2e    = the same metric through the public API from HOST (pinned) arrays: every
         step builds a fit from the host arrays (H2D inside the timed region),
         runs one EM iteration and reads back the energy, lq_F and lq_R.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "edge_subject_em_iterations_per_second"
UNIT = "edge-subject-iterations/s"
H_SUBJ = 500
U_SUBJ = 500
EDGES_PER_GPU = 79800           # Schaefer-400


_REGIONS_OVERRIDE = 0


def regions_for(n_gpus):
    """Smallest N whose edge count reaches n_gpus * 79,800 (400 at one GPU)."""
    if _REGIONS_OVERRIDE > 0:
        return _REGIONS_OVERRIDE
    N = 400
    while N * (N - 1) // 2 < n_gpus * EDGES_PER_GPU:
        N += 1
    return N


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for (name, v) in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(kernel, C, N, H, U):
    """Per-launch algorithmic bytes (DESIGN.md 'Kernels'): 8 bytes per
    edge-patient for every pass over the patient correlations, whatever form the
    kernel actually reads them in."""
    return {
        "K2_estep_qF": 8 * C * U + 16 * C + 48 * C + 16 * N * U,        # bt, S1/S2, lqF+qF out, qR (tiered form: first iteration)
        "K2_estep_qF_coded": 8 * C * U + 16 * C + 48 * C + 16 * N * U,  # the same E-step from the code plane (all later iterations)
        "K2b_region_weights": 8 * C * U + 24 * C + 16 * C * U,          # btT, qF, WT out (two weight differences)
        "K2b_sweep": 16 * C * U + 16 * N * U + 32 * N * U,              # WT once (window overlaps hit L1/L2), qR in, qR/lqR out
        "K2b_sweep_fused": 8 * C * U + 24 * C + 48 * N * U,             # btT once, qF, qR in / out
        "K3b_elm_obj_grad": 8 * C * U + 24 * C + 16 * N * U,            # bt, qF, qR
        "K3b_elm_streams": 8 * C * U + 24 * C + 16 * N * U,             # the same evaluation from the coded plane
        "K3b_elm_const": 8 * C * U + 24 * C + 16 * N * U,
        "K3b_code_plane": C * U + 16 * N * U,                           # one code byte per edge-patient out, q_R states in
        "K4_elm": 8 * C * U + 24 * C + 16 * N * U,
        "K3a_mstep_stats": 24 * C + 16 * N * U,
    }.get(kernel)


# --------------------------------------------------------------------------- CPU arms
def oracle_step_rate(n_regions, steps, warmup, threads=1):
    """The CPU port (oracle/iar_oracle.py, a NumPy restatement of the reference's
    step functions -- the reference itself is Python 2 and /root/reference does
    not travel) on a bounded sample of the workload: the sub-network of the first
    `n_regions` regions x (500 + 500) subjects, consecutive EM iterations."""
    from oracle import iar_oracle as O
    th_true = O.Theta()
    (_, _, _, _, b, bt) = O.sample(th_true, n_regions, H_SUBJ, U_SUBJ, np.random.RandomState(0))
    th = O.Theta()
    th.eta += 0.1
    (lqF, lqR) = O.init_lps(n_regions, U_SUBJ)
    (lpB, pBt, lM) = O.update_lps(b, bt, th)
    C = b.shape[0]
    times, nfev = [], []
    pool = None
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=threads)
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        (lqF, lqR, lM, e, nf) = O.em_iteration(b, bt, th, lqF, lqR, lpB, pBt, lM, pool=pool, chunks=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            nfev.append(nf)
    total = float(sum(times))
    return dict(rate=C * (H_SUBJ + U_SUBJ) * steps / total, ms_per_step=1e3 * total / steps,
                C=C, nfev=float(np.mean(nfev)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_regions = 100
    threads = max(1, min(os.cpu_count() or 1, 32))
    r = oracle_step_rate(n_regions, args.steps, args.warmup, threads)
    sample = ("oracle port (NumPy restatement of fcdiff/fit.py step functions; the (eta, epsilon) objective, "
              "the dominant cost, is split over %d threads, the rest is single-threaded like the reference), "
              "sub-network of %d regions (%d edges) x %d+%d subjects, %d consecutive EM iterations, "
              "%.1f objective evals/iter" % (threads, n_regions, r["C"], H_SUBJ, U_SUBJ, args.steps, r["nfev"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rate"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # `config` names what RAN: a sub-network of the GPU arm's atlas (the oracle needs ~13 GB and ~15 min
        # per EM iteration on the full Schaefer-400 x 500+500 problem; the rate is per edge-subject)
        "config": dict({k: v for (k, v) in workload_config(args.gpus).items() if k != "l2"}, regions=n_regions, edges=r["C"],
                       workload="bounded sample of the GPU arm's workload: sub-network of the first %d regions "
                                "(%d edges) x (%d controls + %d patients); the GPU arm runs %d regions"
                                % (n_regions, r["C"], H_SUBJ, U_SUBJ, regions_for(args.gpus)),
                       parallelism="%d host threads on the (eta, epsilon) objective, 1 elsewhere" % threads),
        "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    N = regions_for(n_gpus)
    return {"workload": "Schaefer-400 x (500 controls + 500 patients), BASELINE.json configs[2]"
                        if (n_gpus == 1 and N == 400 and H_SUBJ == 500) else
                        "%d-region atlas (%d edges over %d GPU) x (%d controls + %d patients)" % (N, N * (N - 1) // 2, n_gpus, H_SUBJ, U_SUBJ),
            "regions": N, "edges": N * (N - 1) // 2, "controls": H_SUBJ, "patients": U_SUBJ,
            "edge_lookup": "reference", "storage": "f64",
            "l2": "per-step working set (responsibility planes, dominant-state planes, WT, coded plane and records: > 2 GB per GPU) "
                  "exceeds the 126 MB L2; no flush needed",
            "parallelism": "edge shards x%d, patient-sharded region sweep" % n_gpus}


# --------------------------------------------------------------------------- GPU arm
def em_step(fit):
    """One pass of the loop body fcdiff/fit.py:76-80."""
    fit._update_lq_F()
    fit._update_lq_R()
    fit._update_theta()
    fit._update_lps()
    return fit._eval_energy()


def sharded_parity(fcdiff, shards, torch):
    """Driver-run correctness of the multi-GPU path: BASELINE.json configs[1] (AAL-90 x 50 + 50) fitted
    for 3 iterations edge-sharded over all ranks and on this rank's GPU alone, from the same inputs."""
    (N, H, U) = (90, 50, 50)
    model = fcdiff.UnsharedRegionModel()
    (_, _, _, _, b, bt) = model.sample_device(N, H, U)
    (b, bt) = (b.cpu().numpy(), bt.cpu().numpy())

    def fit_with(sh):
        f = fcdiff.fit.UnsharedRegionFit()
        f.model = fcdiff.UnsharedRegionModel()
        f.model.eta += 0.1
        (f.b, f.bt) = (b, bt)
        f.shards = sh
        f.max_iters = 3
        f.rel_tol = -1.0
        f.run()
        return f

    (fs, f1) = (fit_with(shards), fit_with(None))
    rel = lambda a, c: float(np.max(np.abs(np.asarray(a) - np.asarray(c)) / np.maximum(np.abs(np.asarray(c)), 1e-300)))
    th = lambda f: [f.model.pi, f.model.eta, f.model.epsilon] + list(np.asarray(f.model.gamma))
    out = {"config": "AAL-90 (4005 edges) x 50 + 50, 3 iterations, sharded over %d ranks vs one GPU" % shards.world,
           "max_rel_energy": rel(fs.energy, f1.energy), "max_rel_theta": rel(th(fs), th(f1)),
           "max_abs_lq_F": float(np.max(np.abs(fs._lq_F - f1._lq_F))),
           "max_abs_lq_R": float(np.max(np.abs(fs._lq_R - f1._lq_R)))}
    out["ok"] = bool(out["max_rel_energy"] < 1e-9 and out["max_rel_theta"] < 1e-9 and out["max_abs_lq_F"] < 1e-6
                     and out["max_abs_lq_R"] < 1e-6)
    t = torch.tensor([0.0 if out["ok"] else 1.0], dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(t)
    out["ok_all_ranks"] = bool(t.item() == 0.0)
    return out


def run_cfg4(fcdiff, _dev, shards, torch, dist, peak):
    """BASELINE.json configs[3]: 1000-region parcellation (499,500 edges) x 2000 subjects (1000 + 1000), full
    fit with free-energy tracking, the SAME problem at every GPU count (strong scaling: edges and the
    sweep's patients are sharded over the ranks).  One warm-up fit, one timed fit of 10 iterations from
    the uniform start."""
    (N, H, U) = (1000, 1000, 1000)
    C = N * (N - 1) // 2
    world = 1 if shards is None else shards.world
    (c0, Cl) = (0, C) if shards is None else shards.span(C)
    need_gb = 8e-9 * Cl * (H + U) + 8e-9 * Cl * U * 7.2 + 8e-9 * C * (U / world) * 6.0
    free_gb = torch.cuda.mem_get_info()[0] / 1e9 + torch.cuda.memory_reserved() / 1e9 - torch.cuda.memory_allocated() / 1e9
    if need_gb > 0.9 * free_gb:
        return {"skipped": "needs ~%.0f GB per GPU, %.0f GB free" % (need_gb, free_gb)}
    torch.cuda.empty_cache()
    model = fcdiff.UnsharedRegionModel()
    (_, _, _, _, b_dev, bt_dev) = model.sample_device(N, H, U, c0=c0, C=Cl)

    def one(profile):
        f = fcdiff.fit.UnsharedRegionFit()
        f.model = fcdiff.UnsharedRegionModel()
        f.model.eta += 0.1
        (f.b, f.bt) = (b_dev, bt_dev)
        if shards is not None:
            f.shards = shards
            f.n_edges = C
        f.max_iters = 10
        f.rel_tol = -1.0
        f.profile = profile
        if shards is not None:
            dist.barrier()
        torch.cuda.synchronize()
        (e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        e0.record()
        f.run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if shards is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return f, float(t.item())

    (f, _) = one(None)
    del f
    timers = _dev.KernelTimers()
    (f, ms) = one(timers)
    ks = timers.summary()
    Ul = U if shards is None else shards.span(U)[1]
    kernels = {}
    for (name, (cnt, total_ms, mean_ms)) in sorted(ks.items()):
        ab = algorithmic_bytes(name, Cl, N, H, Ul if name in ("K2b_region_weights", "K2b_sweep") else U)
        ent = {"launches": cnt, "mean_ms": mean_ms, "share_of_step": total_ms / ms}
        if ab:
            ent["frac_of_hbm_peak"] = ab / (mean_ms * 1e-3) / 1e9 / peak
        kernels[name] = ent
    dom = max((k for k in kernels if "frac_of_hbm_peak" in kernels[k]), key=lambda k: kernels[k]["share_of_step"])
    out = {"workload": "1000 regions (499,500 edges) x (1000 controls + 1000 patients), BASELINE.json configs[3]",
           "scaling": "strong", "n_gpus": world, "iterations": len(f.energy) - 1, "ms_per_step": ms / (len(f.energy) - 1),
           "value": C * (H + U) * (len(f.energy) - 1) / (ms * 1e-3), "unit": UNIT,
           "objective_evals_per_step": float(np.mean(f.n_objective_evals)),
           "energy_trace": [float(e) for e in f.energy], "energy_finite": bool(np.all(np.isfinite(f.energy))),
           "dominant_kernel": dom, "dominant_frac_of_hbm_peak": kernels[dom]["frac_of_hbm_peak"],
           "roofline_traffic": None,       # no ncu --set full capture at this configuration
           "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for (kk, vv) in v.items()}
                       for (k, v) in kernels.items()},
           "hbm_gb_allocated": torch.cuda.max_memory_allocated() / 1e9}
    del f, b_dev, bt_dev
    torch.cuda.empty_cache()
    return out


def run_k1(fcdiff, torch, peaks):
    """The stage in front of the fit in BASELINE.json configs[1], [2] ("1200-TR time series -> correlations ->
    EM fit"; the reference starts at correlations, fcdiff/fit.py:20-23): Schaefer-400 x 1200 TRs x (500 + 500)
    subjects, fp32 time series resident in HBM -> standardise -> split-TF32 Gram on tcgen05 -> Fisher z ->
    edge-major (C, S) fp64 correlations (csrc/fcd_corr*.cu).  Tensor roofline: TF32 dense = half the measured
    bf16 rate; the error-compensated product costs three MMAs, only the lower triangle is needed."""
    from fcdiff_b200 import corr
    (S, N, T) = (H_SUBJ + U_SUBJ, 400, 1200)
    C = N * (N - 1) // 2
    g = torch.Generator(device="cuda").manual_seed(1)
    # planted structure instead of white noise: eight latent signals, every region loads on one of them with
    # +0.55 or -0.27 (same signal: correlations +0.30 / -0.15 / +0.07 -- the model's positive and negative
    # states -- different signals: 0); a patient's region is anomalous with probability 0.05 (loading sign flipped)
    K = 8
    fac = torch.arange(N, device="cuda") % K
    load = torch.where(torch.rand((N,), device="cuda", generator=g) < 0.5, 0.55, -0.27).to(torch.float32)
    a = load[None, :].repeat(S, 1)
    flip = torch.rand((U_SUBJ, N), device="cuda", generator=g) < 0.05
    a[H_SUBJ:][flip] *= -1.0
    ts = torch.randn((S, N, T), dtype=torch.float32, device="cuda", generator=g)
    ts *= torch.sqrt(1.0 - a * a)[:, :, None]
    z = torch.randn((S, K, T), dtype=torch.float32, device="cuda", generator=g)
    ts += a[:, :, None] * z[:, fac, :]
    del z
    out = torch.empty((C, S), dtype=torch.float64, device="cuda")
    for _ in range(2):
        corr.correlations_device(ts, fisher=True, out=out)
    (e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    R = 3
    torch.cuda.synchronize()
    e0.record()
    for _ in range(R):
        corr.correlations_device(ts, fisher=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    useful = 2.0 * S * C * T                                   # one multiply-add per (edge, time point)
    nt = (N + 127) // 128
    tail = N - (nt - 1) * 128
    tailpad = (tail + 15) // 16 * 16
    cols = sum((tailpad if (ti == nt - 1 and tailpad < 128) else 128) for ti in range(nt) for tj in range(ti + 1))
    issued = 3 * 2.0 * S * 128 * cols * ((T + 31) // 32 * 32)  # three MMAs per product over the computed tiles
    tf32_peak = 0.5 * peaks.get("bf16_tflops", 2250.0)
    res = {"workload": "Schaefer-400 x 1200 TRs x %d subjects: time series -> Fisher-z correlations" % S,
           "ms": ms, "subjects_per_s": S / (ms * 1e-3),
           "useful_tflops": useful / (ms * 1e-3) / 1e12, "issued_tf32_tflops": issued / (ms * 1e-3) / 1e12,
           "roofline": {"bound": "tensor", "achieved": issued / (ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": issued / (ms * 1e-3) / 1e12 / tf32_peak,
                        "peak_source": "half the measured dense bf16 rate (MEASURED_PEAKS.json bf16_tflops): TF32",
                        "note": "issued = 3 MMAs (hh, hl, lh) x the tiles computed; useful = the lower triangle once"},
           "includes": "standardise (fp64 mean / norm, TF32 hi / lo planes) + Gram + atanh epilogue"}
    # configs[2] end to end from time series: K1, then one fit.run() on its output (device resident)
    f = fcdiff.fit.UnsharedRegionFit()
    f.model = fcdiff.UnsharedRegionModel()
    (f.b, f.bt) = (out[:, :H_SUBJ].contiguous(), out[:, H_SUBJ:].contiguous())
    f.max_iters = 3
    f.run()                                                     # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    corr.correlations_device(ts, fisher=True, out=out)
    f = fcdiff.fit.UnsharedRegionFit()
    f.model = fcdiff.UnsharedRegionModel()
    (f.b, f.bt) = (out[:, :H_SUBJ].contiguous(), out[:, H_SUBJ:].contiguous())
    f.max_iters = 3
    f.rel_tol = -1.0
    f.run()
    torch.cuda.synchronize()
    res["timeseries_to_fit_ms"] = 1e3 * (time.perf_counter() - t0)
    res["timeseries_to_fit_what"] = ("K1 on %d subjects + fit.run() with 3 EM iterations on its output (series with "
                                     "planted positive / negative correlations and 5 %% anomalous patient regions)" % S)
    res["energy_trace"] = [float(e) for e in f.energy]
    res["energy_finite"] = bool(np.all(np.isfinite(f.energy)))
    del ts, out, f
    torch.cuda.empty_cache()
    return res


def run_cfg5(fcdiff, shards, torch, dist, n_replicas):
    """BASELINE.json configs[4]: Schaefer-400 x 1000 subjects, `n_replicas`-way group-label permutation
    sweep.  The responsibility planes are built once for all 1000 subjects; a replica selects its
    columns (fcdiff_b200/sweep.py: SharedPlanes); replicas are sharded over the ranks, nothing is
    exchanged.  Every replica is a plain ``fit.run()`` with the reference's defaults."""
    from fcdiff_b200 import sweep
    (N, H, U) = (400, 500, 500)
    C = N * (N - 1) // 2
    (rank, world) = (0, 1) if shards is None else (shards.rank, shards.world)
    model = fcdiff.UnsharedRegionModel()
    (_, _, _, _, b, bt) = model.sample_device(N, H, U)            # the same matrix on every rank (same key)
    corr = torch.cat([b, bt], dim=1)
    labels = np.r_[np.zeros(H, bool), np.ones(U, bool)]
    start = fcdiff.UnsharedRegionModel()
    start.eta += 0.1

    def sync():
        if shards is not None:
            dist.barrier()
        torch.cuda.synchronize()

    streams = sweep.default_streams(world)              # replicas in flight per GPU (2 when the host has the cores)
    sweep.permutation_sweep(corr, labels, 4 * streams * world - 1, model=start, rank=rank, world=world, gather=False,
                            streams=streams)              # warm-up: every stream's buffers and per-stream state
    sync()
    t0 = time.perf_counter()
    res = sweep.permutation_sweep(corr, labels, n_replicas - 1, model=start, rank=rank, world=world, gather=False,
                                  streams=streams)
    sync()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    iters = torch.tensor([float(sum(r["iterations"] for r in res.values())), float(len(res))], dtype=torch.float64,
                         device="cuda")
    if shards is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(iters)
    out = {"workload": "Schaefer-400 (79,800 edges) x 1000 subjects, %d relabelings (replica 0 = observed labels), "
                       "BASELINE.json configs[4]" % n_replicas,
           "replicas": int(iters[1].item()), "n_gpus": world, "scaling": "strong (replicas sharded over ranks)",
           "seconds": float(dt.item()), "replicas_per_second": float(iters[1].item() / dt.item()),
           "em_iterations": int(iters[0].item()),
           "value": float(C * (H + U) * iters[0].item() / dt.item()), "unit": UNIT,
           "replicas_in_flight_per_gpu": streams,
           "fit": "fit.run(), max_iters 10, rel_tol 1e-5, the reference's convergence rule; planes built once, "
                  "per-replica column selection"}
    if rank == 0:                                  # replica 0 is the plain fit of (b, bt)
        f = fcdiff.fit.UnsharedRegionFit()
        f.model = fcdiff.UnsharedRegionModel()
        f.model.eta += 0.1
        (f.b, f.bt) = (b, bt)
        f.run()
        e0 = np.asarray(res[0]["energy"])
        out["replica0_vs_plain_fit_max_rel"] = float(np.max(np.abs(e0 - np.asarray(f.energy)) / np.abs(np.asarray(f.energy)))) \
            if len(e0) == len(f.energy) else None
        out["energy_final_observed_vs_permuted_mean"] = [float(e0[-1]), float(np.mean([r["energy"][-1] for (i, r) in res.items() if i > 0]))]
        del f
    del corr, b, bt, res
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.build()
    import fcdiff_b200 as fcdiff
    from fcdiff_b200 import _dev, _lib
    from fcdiff_b200 import dist as fdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    shards = fdist.init_from_env("nccl") if world > 1 else None
    lib = _lib.load()

    N = regions_for(world)
    C = N * (N - 1) // 2
    (H, U) = (H_SUBJ, U_SUBJ)
    (c0, Cl) = (0, C) if shards is None else shards.span(C)

    # synthetic inputs drawn on the device by the Philox sampler (K5): every rank
    # draws its own edge rows from the same key (the sampler is shard-invariant)
    true_model = fcdiff.UnsharedRegionModel()
    (_, _, _, _, b_dev, bt_dev) = true_model.sample_device(N, H, U, c0=c0, C=Cl)
    torch.cuda.synchronize()

    def new_fit(b, bt):
        fit = fcdiff.fit.UnsharedRegionFit()
        fit.model = fcdiff.UnsharedRegionModel()
        fit.model.eta += 0.1
        fit.b, fit.bt = b, bt
        if shards is not None:
            fit.shards = shards
            fit.n_edges = C
        if os.environ.get("FCD_FUSED_SWEEP"):                 # manual experiments
            fit.fused_sweep = os.environ["FCD_FUSED_SWEEP"] == "1"
        return fit

    def barrier():
        if shards is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ITERS_PER_FIT = 10                         # the reference's default max_iters (fcdiff/fit.py:36)

    def run_fits(n_steps, profile=None):
        """n_steps EM iterations as fresh fits from the uniform start, each a plain ``fit.run()``
        with max_iters = 10 (the last one shorter) and the convergence test disabled."""
        (energies, nfev) = ([], [])
        left = n_steps
        while left > 0:
            f = new_fit(b_dev, bt_dev)
            f.max_iters = min(ITERS_PER_FIT, left)
            f.rel_tol = -1.0
            f.profile = profile
            f.run()
            left -= f.max_iters
            energies.append([float(e) for e in f.energy])
            nfev.extend(f.n_objective_evals)
            del f
        return energies, nfev

    # ---- headline: fresh fits from the uniform start, device-resident inputs
    run_fits(max(args.warmup, 3))              # warm-up: first-use allocations, log table, peer windows
    timers = _dev.KernelTimers()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    lib.fcd_launch_count_reset()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    (fit_energies, nfev_list) = run_fits(args.steps, timers)
    ev1.record()
    barrier()
    launches = int(lib.fcd_launch_count())
    clk = clocks.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    ksum = timers.summary()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if shards is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = C * (H + U) * args.steps / (ms * 1e-3)
    nfev = float(np.mean(nfev_list)) if nfev_list else 0.0
    energies = fit_energies[0]

    # ---- steady state: consecutive iterations of one fit after warm-up iterations (round 1's `value`)
    fit = new_fit(b_dev, bt_dev)
    fit._init_lps(N, H, U)
    fit._update_lps()
    fit._eval_energy()
    for _ in range(max(args.warmup, 5)):
        em_step(fit)
    fit.n_objective_evals = []
    ss_steps = max(5, min(args.steps, 20))
    barrier()
    es0 = torch.cuda.Event(enable_timing=True)
    es1 = torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(ss_steps):
        em_step(fit)
    es1.record()
    barrier()
    tss = torch.tensor([es0.elapsed_time(es1)], dtype=torch.float64, device="cuda")
    if shards is not None:
        dist.all_reduce(tss, op=dist.ReduceOp.MAX)
    steady = {"value": C * (H + U) * ss_steps / (float(tss.item()) * 1e-3), "unit": UNIT,
              "ms_per_step": float(tss.item()) / ss_steps, "steps": ss_steps,
              "objective_evals_per_step": float(np.mean(fit.n_objective_evals)) if fit.n_objective_evals else 0.0,
              "what": "consecutive iterations of one fit after %d warm-up iterations (posteriors settled)"
                      % max(args.warmup, 5)}
    del fit                                    # release its cache planes before the next fits allocate theirs

    # ---- time to converge (north_star target): whole fit from device-resident inputs
    def timed_fit(**opts):
        f = new_fit(b_dev, bt_dev)
        f.max_iters = 100
        for (k, v) in opts.items():
            setattr(f, k, v)
        barrier()
        t0 = time.perf_counter()
        f.run()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if shards is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out = {"seconds": float(tt.item()), "iterations": len(f.energy) - 1,
               "objective_evals": int(sum(f.n_objective_evals)),
               "energy_first_last": [float(f.energy[0]), float(f.energy[-1])],
               "energy_monotone": bool(np.all(np.diff(f.energy) <= 0))}
        del f
        return out

    converge = timed_fit()
    converge.update({"rel_tol": 1e-5,
                     "what": "fit.run() to the reference's convergence rule (fit.py:138-140) incl. set-up "
                             "(healthy stats, responsibility planes, dominant-state planes)"})
    converge["magnitude_rule"] = dict(timed_fit(convergence_rule="magnitude"),
                                      what="same with (e - e*)/|e| < rel_tol (the reference's rule stops at the first "
                                           "decrease of a negative energy)")
    converge["symmetric_magnitude"] = dict(
        timed_fit(convergence_rule="magnitude", edge_lookup="symmetric"),
        what="edge_lookup='symmetric' (the mathematically intended edge of fit.py:185-186, SURVEY 0.3) with the "
             "magnitude rule: with this lookup every update is a coordinate descent step, so the energy falls "
             "monotonically and the stopping point is a fixed point")

    # ---- N > 1: the sharded fit against the single-device fit on the same inputs (config 2)
    parity = None
    if shards is not None:
        parity = sharded_parity(fcdiff, shards, torch)

    # ---- end-to-end arm: host (pinned) arrays through the public API every step
    e2e = None
    b_host = torch.empty((Cl, H), dtype=torch.float64).pin_memory()
    bt_host = torch.empty((Cl, U), dtype=torch.float64).pin_memory()
    b_host.copy_(b_dev)
    bt_host.copy_(bt_dev)
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        # shard-local host rows -> device (H2D), one EM iteration, results -> host (D2H)
        f = new_fit(b_host.to("cuda", non_blocking=True), bt_host.to("cuda", non_blocking=True)) \
            if shards is not None else new_fit(b_host.numpy(), bt_host.numpy())
        f.max_iters = 1
        f.run()
        return f.energy[-1], f._lq_F, f._lq_R

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        (e_last, lqF_h, lqR_h) = e2e_step()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if shards is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    e2e = {"value": C * (H + U) * e2e_steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": int(8 * Cl * (H + U)),
           "d2h_bytes_per_step": int(8 * (1 + lqF_h.size + lqR_h.size)),
           "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
           "what": "fit.run(max_iters=1) from pinned host arrays: H2D of b, bt + healthy stats + responsibility "
                   "planes + initial energy + one EM iteration + energy; D2H of energy, lq_F, lq_R"}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"

    # ---- K1 (time series -> correlations), one GPU: the stage in front of the fit in configs[1], [2]
    k1 = None
    if world == 1 and not args.no_k1:
        b_host = bt_host = b_dev = bt_dev = None
        torch.cuda.empty_cache()
        try:
            k1 = run_k1(fcdiff, torch, json.load(open(peaks_path)) if os.path.isfile(peaks_path) else {})
        except Exception as exc:
            k1 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- configs[3] at this GPU count (strong scaling), all ranks take part
    cfg4 = None
    if not args.no_cfg4:
        b_host = bt_host = b_dev = bt_dev = None         # release the main workload's inputs
        try:
            cfg4 = run_cfg4(fcdiff, _dev, shards, torch, dist, peak)
        except Exception as exc:                   # an extra block must never cost the headline line
            cfg4 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- configs[4]: the permutation sweep, replicas sharded over the ranks
    cfg5 = None
    if args.replicas > 0:
        try:
            cfg5 = run_cfg5(fcdiff, shards, torch, dist, args.replicas)
        except Exception as exc:
            cfg5 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel (by time inside the timed region)
    kernels = {}
    for (name, (cnt, total_ms, mean_ms)) in sorted(ksum.items()):
        ab = algorithmic_bytes(name, Cl, N, H, U if name != "K2b_region_weights" and name != "K2b_sweep"
                               else (U if shards is None else shards.span(U)[1]))
        ent = {"launches": cnt, "total_ms": total_ms, "mean_ms": mean_ms, "share_of_step": total_ms / ms}
        lm = sorted(timers.launch_ms(name))
        if lm and len(lm) == cnt:                     # one bracket per launch: the spread over the fits' iterations
            ent["min_ms"], ent["median_ms"], ent["max_ms"] = lm[0], lm[len(lm) // 2], lm[-1]
        if ab:
            ent["algorithmic_bytes"] = int(ab)
            ent["achieved_gbs"] = ab / (mean_ms * 1e-3) / 1e9
            ent["frac_of_hbm_peak"] = ent["achieved_gbs"] / peak
        kernels[name] = ent
    dom = max((k for k in kernels if "achieved_gbs" in kernels[k]), key=lambda k: kernels[k]["total_ms"])
    # measured DRAM bytes per launch (ncu --set full) exist for ONE configuration; any other reports null
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        tc = tj.get("_config", {})
        if (tc.get("regions"), tc.get("controls"), tc.get("patients"), tc.get("n_gpus")) == (N, H, U, world):
            traffic = tj.get(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic,
                "peak_source": peak_src,
                "note": "achieved = algorithmic bytes (8 B per edge-patient, SURVEY 8d) / mean launch time; "
                        "see DESIGN.md 'Kernels' for what each kernel actually reads"}

    # the E-step kernel north_star sets its roofline target for (K2: coded form in every iteration but the first)
    roofline_estep = None
    for k2 in ("K2_estep_qF_coded", "K2_estep_qF"):
        if k2 in kernels and "achieved_gbs" in kernels[k2]:
            roofline_estep = {"kernel": k2, "bound": "hbm", "achieved": kernels[k2]["achieved_gbs"], "peak": peak,
                              "unit": "GB/s", "frac": kernels[k2]["frac_of_hbm_peak"],
                              "mean_ms": kernels[k2]["mean_ms"], "launches": kernels[k2]["launches"]}
            break

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = max(1, min(os.cpu_count() or 1, 32))
        r = oracle_step_rate(100, 2, 1, threads)
        cpu = {"value": r["rate"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle port (objective split over %d threads), sub-network of 100 regions (%d edges) x "
                         "500+500 subjects, two EM iterations after one warm-up iteration, %.1f objective evals each"
                         % (threads, r["C"], r["nfev"]),
               "host_cores": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(world), objective_evals_per_step=nfev, iterations_per_fit=ITERS_PER_FIT,
                       start="uniform posteriors (fit.py:84-102), theta = model defaults with eta + 0.1",
                       theta_solver="newton (device-resident)"),
        "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roofline, "roofline_estep": roofline_estep,
        "cpu_baseline": cpu, "steady_state": steady, "time_to_converge": converge, "kernels": kernels,
        "kernel_share_of_step": float(sum(k["total_ms"] for k in kernels.values()) / ms),
        "parity": parity, "k1": k1, "cfg4": cfg4, "cfg5": cfg5,
        "energy_trace": [float(e) for e in energies],
    }
    _emit(json.dumps(line))
    return 0


_emit = print


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-k1", action="store_true", help="skip the K1 (time series -> correlations) block")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the configs[3] (1000 regions x 2000 subjects) block")
    ap.add_argument("--replicas", type=int, default=1000,
                    help="relabelings of the configs[4] permutation sweep block (0 = skip)")
    ap.add_argument("--regions", type=int, default=0, help="override the atlas size (manual experiments)")
    ap.add_argument("--subjects", type=int, default=0, help="override controls = patients (manual experiments)")
    args = ap.parse_args()
    global H_SUBJ, U_SUBJ, _REGIONS_OVERRIDE
    if args.subjects > 0:
        H_SUBJ = U_SUBJ = args.subjects
    _REGIONS_OVERRIDE = args.regions
    if args.impl == "reference":
        return run_reference_arm(args)
    # Exactly one JSON line must reach stdout: libraries (NCCL's version banner, for one)
    # print there too, so stdout is pointed at stderr for the duration of the run and the
    # JSON line is written to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    _emit = lambda text: os.write(real_stdout, (text + "\n").encode())
    rc = run_gpu_arm(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
