#!/usr/bin/env python
"""K1 only (for ncu): one tensor-core correlation pass at a slice of config 3."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcdiff_b200 import corr          # noqa: E402

ts = torch.randn((64, 400, 1200), dtype=torch.float32, device="cuda")
for _ in range(2):
    out = corr.correlations_device(ts, fisher=True)
torch.cuda.synchronize()
print("ok", out.shape)
