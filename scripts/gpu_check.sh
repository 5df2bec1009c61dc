#!/bin/bash
# GPU-box check: parity tests + the bench line with a per-kernel summary (run under gpurun).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [pytest -k expression]'
mkdir -p gpurun_out
KEXPR=${1:-}
if [ -n "$KEXPR" ]; then
    python -m pytest tests -m gpu -q -k "$KEXPR" 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
else
    python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
fi
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json"))
print("value %.3g  ms/step %.3f  steady %.3g (%.3f ms, %.1f evals)  e2e %.3g (%.1f ms)  nfev %.1f launches %d kernel share %.2f"%(
    d["value"], d["ms_per_step"], d["steady_state"]["value"], d["steady_state"]["ms_per_step"],
    d["steady_state"]["objective_evals_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["config"]["objective_evals_per_step"], d["gpu_launches"], d["kernel_share_of_step"]))
for k,v in d["kernels"].items(): print("%-22s %4d  %.3f ms  share %.3f  frac %.3f"%(k, v["launches"], v["mean_ms"], v["share_of_step"], v.get("frac_of_hbm_peak",0)))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "estep", d["roofline_estep"])
print("converge", json.dumps(d["time_to_converge"])[:900])
print("cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
print("energies", d["energy_trace"])
c=d.get("cfg4") or {}; print("cfg4", {k:v for k,v in c.items() if k not in ("kernels","energy_trace")})
print("cfg5", d.get("cfg5"))
print("k1", d.get("k1"))
PY
