"""
TEST INFRASTRUCTURE ONLY -- golden vectors of the BENCHMARKED configuration
(BASELINE.json configs[2]: Schaefer-400 = 79,800 edges x 500 controls + 500 patients).

The oracle (oracle/iar_oracle.py, pinned to the reference at configs 1-2) needs
~13 GB and ~10 minutes per EM iteration at this size -- too slow to run inside the
GPU tests -- so its result for two full ``run()`` iterations from the uniform start is
computed ONCE here, in the build container, and committed as a small fixture:

    python oracle/make_golden_cfg3.py lbfgsb     # SciPy L-BFGS-B, default tolerances (the reference's call)
    python oracle/make_golden_cfg3.py polished   # the same, polished to the minimiser (oracle of the Newton solver)

-> tests/golden/cfg3_run2_<variant>.npz: the energy trace, theta after every iteration,
the MAP labels of q_F and q_R (bit-packed), checksums of the posteriors and the values
of a fixed random sample of 40,000 entries of lq_F and of lq_R.  The inputs are NOT
stored: ``inputs()`` regenerates them bit-identically from NumPy's legacy RandomState
(the same call the GPU test makes).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import iar_oracle as O          # noqa: E402

(N, H, U) = (400, 500, 500)
ITERS = 2
N_SAMPLE = 40000


def inputs():
    """(b, bt) of the golden run: the oracle's sampler at the model defaults, seed 0."""
    (_, _, _, _, b, bt) = O.sample(O.Theta(), N, H, U, np.random.RandomState(0))
    return b, bt


def start_theta():
    th = O.Theta()
    th.eta += 0.1
    return th


def sample_indices(size):
    return np.sort(np.random.RandomState(1).choice(size, N_SAMPLE, replace=False))


def summarise(lq_F, lq_R):
    """What the fixture keeps of the final posteriors (and what the GPU test recomputes from its own)."""
    (qF, qR) = (np.exp(lq_F), np.exp(lq_R))
    return dict(
        map_F=np.packbits(np.argmax(lq_F[:, 0, :], axis=1).astype(np.uint8)[:, None] >> np.arange(2)[None, :] & 1),
        map_R=np.packbits(lq_R[:, :, 1] > lq_R[:, :, 0]),
        gap_F=np.packbits((np.sort(lq_F[:, 0, :], axis=1)[:, 2] - np.sort(lq_F[:, 0, :], axis=1)[:, 1]) > 1e-6),
        gap_R=np.packbits(np.abs(lq_R[:, :, 1] - lq_R[:, :, 0]) > 1e-6),
        sum_qF=qF.sum(axis=(0, 1)), sum_qR=qR.sum(axis=(0, 1)),
        ent_F=float(np.sum(qF * lq_F)), ent_R=float(np.sum(qR * lq_R)),
        lq_F_sample=lq_F.reshape(-1)[sample_indices(lq_F.size)],
        lq_R_sample=lq_R.reshape(-1)[sample_indices(lq_R.size)])


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "polished"
    assert variant in ("lbfgsb", "polished")
    (b, bt) = inputs()
    th = start_theta()
    trace = dict(pi=[], eta=[], epsilon=[], gamma=[], nfev=[], seconds=[])
    t0 = time.time()

    def record(i, th, lq_F, lq_R, e, nfev):
        trace["pi"].append(th.pi)
        trace["eta"].append(th.eta)
        trace["epsilon"].append(th.epsilon)
        trace["gamma"].append(np.array(th.gamma))
        trace["nfev"].append(nfev)
        trace["seconds"].append(time.time() - t0)
        print("iteration", i, "energy", e, "theta", th.pi, th.eta, th.epsilon, "nfev", nfev, "t", time.time() - t0, flush=True)

    out = O.run(b, bt, th, max_iters=ITERS, rel_tol=-1.0, record=record, polish=(variant == "polished"))
    fix = summarise(out["lq_F"], out["lq_R"])
    fix.update(energy=np.array(out["energy"]), **{k: np.array(v) for (k, v) in trace.items()})
    fix["shape"] = np.array([N, H, U, ITERS])
    path = os.path.join(ROOT, "tests", "golden", "cfg3_run2_%s.npz" % variant)
    np.savez_compressed(path, **fix)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
