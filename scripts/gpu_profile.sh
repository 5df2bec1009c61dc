#!/bin/bash
# Profile collection (run under gpurun on one B200): plain run first, then the ncu launch list and one
# --set full capture of the hot kernels of the same command.  Outputs land in gpurun_out/.
#   bash scripts/gpu_profile.sh TAG [kernel regex] [skip] [count]
set -o pipefail
mkdir -p gpurun_out
TAG=${1:-r02a}
REGEX=${2:-"elm_coded|estep_qF_coded|sweep_blocked|sweep_kernel|region_weights|code_plane|record_keys|record_weights|record_half|row_logsums"}
SKIP=${3:-60}
COUNT=${4:-14}
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cfg4 --no-k1 --replicas 0"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err || { echo "plain run failed"; tail -5 gpurun_out/plain_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
