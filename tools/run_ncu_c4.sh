#!/bin/bash
# ncu --set full of the tiny-cloud kernels (one cloud per read pair): one 10 M-pair batch of c4.
# Usage (one GPU): tools/gpurun_retry.sh 1500 1 'bash tools/run_ncu_c4.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ARGS="--config c4 --pairs 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 python bench.py $ARGS > gpurun_out/plain_c4.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_c4.log; exit 1; }
tail -1 gpurun_out/plain_c4.log | cut -c1-200
# warm-up step = 9 launches (tnf, 3 x lookup, 3 x collect, 2 x normalize) of these kernels; profile the timed step's
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'normalize_rows|bucket_collect|bucket_lookup|tnf_kernel' --launch-skip 9 --launch-count 9 \
    -o gpurun_out/prof_r02_c4 -f python bench.py $ARGS > gpurun_out/ncu_c4.log 2>&1
echo "ncu exit $?"
ls -la gpurun_out/prof_r02_c4.ncu-rep
