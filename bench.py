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
(K3a), the L-BFGS-B solve for (eta, epsilon): one code pass, then J
objective+gradient evaluations over the coded dominant-state plane (K3b), free energy (K4).  Steps are consecutive
iterations of one fit.

value  = C*(H+U)*K / t with inputs resident in HBM (CUDA events, max over ranks)
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
        "K2_estep_qF": 8 * C * U + 16 * C + 48 * C + 16 * N * U,        # bt, S1/S2, lqF+qF out, qR
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
        "config": workload_config(args.gpus),
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
            "l2": "per-step working set (responsibility planes, patient-major planes, WT, coded plane and records: > 2 GB per GPU) "
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

    # ---- device-resident arm
    fit = new_fit(b_dev, bt_dev)
    fit._init_lps(N, H, U)
    fit._update_lps()
    energies = [fit._eval_energy()]
    for _ in range(args.warmup):
        energies.append(em_step(fit))
    timers = _dev.KernelTimers()
    fit.profile = timers
    fit.n_objective_evals = []
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    lib.fcd_launch_count_reset()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        energies.append(em_step(fit))
    ev1.record()
    barrier()
    launches = int(lib.fcd_launch_count())
    clk = clocks.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    ksum = timers.summary()
    fit.profile = None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if shards is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = C * (H + U) * args.steps / (ms * 1e-3)
    nfev = float(np.mean(fit.n_objective_evals)) if fit.n_objective_evals else 0.0
    del fit                                    # release its cache planes before the next fits allocate theirs

    # ---- time to converge (north_star target): whole fit from device-resident inputs
    barrier()
    fitc = new_fit(b_dev, bt_dev)
    fitc.max_iters = 100
    tc0 = time.perf_counter()
    fitc.run()
    barrier()
    t_conv = time.perf_counter() - tc0
    tconv = torch.tensor([t_conv], dtype=torch.float64, device="cuda")
    if shards is not None:
        dist.all_reduce(tconv, op=dist.ReduceOp.MAX)
    converge = {"seconds": float(tconv.item()), "iterations": len(fitc.energy) - 1,
                "rel_tol": fitc.rel_tol, "objective_evals": int(sum(fitc.n_objective_evals)),
                "what": "fit.run() to the reference's convergence rule (fit.py:138-140) incl. set-up "
                        "(healthy stats, Gaussian cache, patient-major planes)"}
    del fitc
    fitm = new_fit(b_dev, bt_dev)
    fitm.max_iters = 100
    fitm.convergence_rule = "magnitude"
    barrier()
    tm0 = time.perf_counter()
    fitm.run()
    barrier()
    tmag = torch.tensor([time.perf_counter() - tm0], dtype=torch.float64, device="cuda")
    if shards is not None:
        dist.all_reduce(tmag, op=dist.ReduceOp.MAX)
    converge["magnitude_rule"] = {"seconds": float(tmag.item()), "iterations": len(fitm.energy) - 1,
                                  "objective_evals": int(sum(fitm.n_objective_evals)),
                                  "what": "same with (e - e*)/|e| < rel_tol (the reference's rule stops at the first "
                                          "decrease of a negative energy)"}
    del fitm

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
           "what": "fit.run(max_iters=1) from pinned host arrays: H2D of b, bt + healthy stats + patient-major "
                   "copy + initial energy + one EM iteration + energy; D2H of energy, lq_F, lq_R"}

    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel (by time inside the timed region)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    kernels = {}
    for (name, (cnt, total_ms, mean_ms)) in sorted(ksum.items()):
        ab = algorithmic_bytes(name, Cl, N, H, U if name != "K2b_region_weights" and name != "K2b_sweep"
                               else (U if shards is None else shards.span(U)[1]))
        ent = {"launches": cnt, "total_ms": total_ms, "mean_ms": mean_ms, "share_of_step": total_ms / ms}
        if ab:
            ent["algorithmic_bytes"] = int(ab)
            ent["achieved_gbs"] = ab / (mean_ms * 1e-3) / 1e9
            ent["frac_of_hbm_peak"] = ent["achieved_gbs"] / peak
        kernels[name] = ent
    dom = max((k for k in kernels if "achieved_gbs" in kernels[k]), key=lambda k: kernels[k]["total_ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic,
                "peak_source": peak_src,
                "note": "achieved = algorithmic bytes (8 B per edge-patient, SURVEY 8d) / mean launch time; "
                        "see DESIGN.md 'Kernels' for what each kernel actually reads"}

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
        "config": dict(workload_config(world), objective_evals_per_step=nfev),
        "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roofline,
        "cpu_baseline": cpu, "time_to_converge": converge, "kernels": kernels,
        "energy_trace": [float(e) for e in energies[:4]] + ["..."] + [float(energies[-1])],
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
