#!/bin/bash
# Row-group E-step kernel: warps per CTA (FCD_K2R_WARPS) and the warp-per-row kernel (FCD_K2=warp), same bench command.
run() {
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cfg4 --no-k1 --replicas 0 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
k=d['kernels']['K2_estep_qF_coded']
print('$1: K2 mean %.4f ms frac %.3f  ms/step %.3f steady %.3f'%(k['mean_ms'],k['frac_of_hbm_peak'],d['ms_per_step'],d['steady_state']['ms_per_step']))"
}
run "warp-per-row (default)"
for w in 8 12 16; do FCD_K2=rows FCD_K2R_WARPS=$w run "rows, $w warps"; done
