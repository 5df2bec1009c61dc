#!/bin/bash
# Multi-GPU check (gpurun --gpus N): sharded = single-device fit, then the bench line at N GPUs.
NG=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $RUN tests/multi_gpu_check.py > gpurun_out/dist_check_g$NG.log 2>&1; echo "dist_check rc=$?"; grep -E "OK|MISMATCH|DIST_CHECK|Error|error" gpurun_out/dist_check_g$NG.log | tail -20
timeout 900 $RUN bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_g$NG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_g$NG.json"))
print("value %.3g  ms/step %.3f  steady %.3g (%.3f ms)  e2e %.3g  nfev %.1f kernel share %.2f"%(
    d["value"], d["ms_per_step"], d["steady_state"]["value"], d["steady_state"]["ms_per_step"], d["e2e"]["value"],
    d["config"]["objective_evals_per_step"], d["kernel_share_of_step"]))
for k,v in d["kernels"].items(): print("%-22s %4d  %.3f ms  share %.3f  frac %.3f"%(k, v["launches"], v["mean_ms"], v["share_of_step"], v.get("frac_of_hbm_peak",0)))
print("parity", d["parity"])
c=d["cfg4"]; print("cfg4", {k:v for k,v in c.items() if k!="kernels"})
PY
