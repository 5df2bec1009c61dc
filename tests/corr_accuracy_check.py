#!/usr/bin/env python
"""K1 check on the GPU box: tensor-core Gram vs the SIMT fp64-accumulate Gram vs numpy, plus timing."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcdiff_b200 import corr, _dev          # noqa: E402
from oracle import iar_oracle as O          # noqa: E402


def run(ts_dev, fisher, simt):
    if simt:
        os.environ["FCD_CORR_SIMT"] = "1"
    else:
        os.environ.pop("FCD_CORR_SIMT", None)
    out = corr.correlations_device(ts_dev, fisher=fisher)
    torch.cuda.synchronize()
    return out


def main():
    rng = np.random.RandomState(0)
    for (S, N, T) in [(4, 128, 256), (3, 90, 200), (5, 130, 64), (9, 400, 1200)]:
        mix = rng.standard_normal((N, N)) * 0.3 + np.eye(N)
        ts = np.einsum("nm,smt->snt", mix, rng.standard_normal((S, N, T))).astype(np.float32)
        ts_dev = _dev.upload(ts, np.float32)
        r_tc = _dev.download(run(ts_dev, False, False))
        r_si = _dev.download(run(ts_dev, False, True))
        r_np = O.corr_fisherz(ts[: min(S, 3)], fisher=False)
        print("S=%d N=%d T=%d  |tc - simt| max %.3e   |simt - numpy| max %.3e   |tc - numpy| max %.3e"
              % (S, N, T, np.abs(r_tc - r_si).max(), np.abs(r_si[:, :r_np.shape[1]] - r_np).max(),
                 np.abs(r_tc[:, :r_np.shape[1]] - r_np).max()), flush=True)
    # timing at config-3 shape (a slice of the subjects)
    (S, N, T) = (200, 400, 1200)
    ts_dev = torch.randn((S, N, T), dtype=torch.float32, device="cuda")
    for simt in (False, True):
        run(ts_dev, True, simt)
        t0 = time.perf_counter()
        for _ in range(3):
            run(ts_dev, True, simt)
        dt = (time.perf_counter() - t0) / 3
        print("%s: S=%d N=%d T=%d  %.2f ms  (%.1f subjects/ms)" % ("simt" if simt else "tc  ", S, N, T, dt * 1e3, S / (dt * 1e3)))


if __name__ == "__main__":
    main()
