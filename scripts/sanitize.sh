#!/bin/bash
# compute-sanitizer passes over the hot path (SURVEY 5 "race detection / sanitizers"; the reference
# has none).  Run on a GPU box:   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
#
#   memcheck   out-of-bounds / misaligned accesses of every kernel the small fits launch
#   racecheck  shared-memory hazards: the per-warp TMA rings + mbarriers of the streaming kernels,
#              the sweep's shared q_R rows, the deterministic last-CTA reductions
#   synccheck  barrier misuse (mbarrier / __syncwarp / __syncthreads in divergent code)
#   initcheck  reads of uninitialised device memory (global)
#
# Workload: scripts/sanitize_workload.py = configs 1-2 of BASELINE.json through the public API
# (fit.run, both edge lookups, coded and tiered K3b paths, the mu/sigma step, the sampler, K1), sized so
# that each tool finishes in a few minutes (the tools slow kernels down 10-100x).  With two GPUs
# visible (gpurun --gpus 2) the 2-rank peer-window exchange is checked too.
# Summaries land in gpurun_out/sanitize_<tool>.log; copy them to profiles/ to have them judged.
set -u
mkdir -p gpurun_out
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
WORK="python scripts/sanitize_workload.py"
rc=0
for tool in memcheck racecheck synccheck initcheck; do
    extra=""
    [ "$tool" = "memcheck" ] && extra="--leak-check no"
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    # torch's caching allocator hands out uninitialised blocks by design; only OUR kernels are of interest
    timeout 1200 $SAN --tool $tool $extra --kernel-regex kns=fcd --print-limit 20 --error-exitcode 3 \
        $WORK ${SANITIZE_ARGS:-} > gpurun_out/sanitize_$tool.log 2>&1
    st=$?
    echo "== $tool: exit $st; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
    [ $st -ne 0 ] && rc=$st
done
NG=$(nvidia-smi -L 2>/dev/null | wc -l)
if [ "$NG" -ge 2 ]; then
    # the peer-window exchange (csrc/fcd_comm.cu) between two ranks, each under memcheck
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 --no-python $SAN --tool memcheck --leak-check no --kernel-regex kns=fcd \
        --error-exitcode 3 python scripts/sanitize_workload.py --dist > gpurun_out/sanitize_dist_memcheck.log 2>&1
    st=$?
    echo "== 2-rank memcheck: exit $st; $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_dist_memcheck.log | tail -2 | tr '\n' ' ')"
    [ $st -ne 0 ] && rc=$st
fi
exit $rc
