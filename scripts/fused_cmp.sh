#!/bin/bash
# Region sweep with the weight tensor (region_weights + sweep over WT) against the fused form, same bench command.
for f in 0 1; do
FCD_FUSED_SWEEP=$f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cfg4 --no-k1 --replicas 0 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
k=d['kernels']
names=[n for n in k if n.startswith('K2b')]
print('fused_sweep=$f: ms/step %.3f steady %.3f  '%(d['ms_per_step'],d['steady_state']['ms_per_step'])+'  '.join('%s %.3f ms'%(n,k[n]['mean_ms']) for n in names), ' energy', d['energy_trace'][-1])"
done
