#!/bin/bash
# GPU-box check: parity tests + a short bench with a per-kernel summary (run under gpurun).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench3.json"))
print("value %.3g  ms/step %.2f e2e %.3g (%.1f ms) nfev %.1f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["objective_evals_per_step"], d["gpu_launches"]))
for k,v in d["kernels"].items(): print(k, v["launches"], "%.3f ms"%v["mean_ms"], "share %.2f"%v["share_of_step"], "frac %.3f"%v.get("frac_of_hbm_peak",0))
print(d["clocks"])
PY
