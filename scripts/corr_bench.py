#!/usr/bin/env python
"""K1 only (for ncu and quick timing): tensor-core correlation passes at a slice of config 3.
    python scripts/corr_bench.py [subjects] [regions] [time points]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcdiff_b200 import corr          # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = int(sys.argv[2]) if len(sys.argv) > 2 else 400
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
ts = torch.randn((S, N, T), dtype=torch.float32, device="cuda")
out = corr.correlations_device(ts, fisher=True)
(e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    corr.correlations_device(ts, fisher=True, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("S=%d N=%d T=%d: %.3f ms per pass (standardise + Gram), %.0f subjects/ms, useful %.1f TFLOP/s"
      % (S, N, T, ms, S / ms, 2.0 * S * (N * (N - 1) // 2) * T / (ms * 1e-3) / 1e12))
