#!/bin/bash
# parity tests + bench, then one ncu --set full capture (with source) of the K3b kernels of the same bench command
mkdir -p gpurun_out
bash scripts/gpu_check.sh || exit 1
NAME=${1:-prof_r01l}
ncu --set full --clock-control none --import-source on \
    -k regex:"elm_coded|code_plane|record_fill|bucket_records|record_scan|pstar_edge" -s 6 -c 14 \
    -o gpurun_out/$NAME -f python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
