# usage: bash tools/run_ncu.sh <tag> [pairs]  - plain run first, then one ncu --set full capture of the sliced kernels
TAG=${1:-r01}; PAIRS=${2:-5000000}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --pairs $PAIRS --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'bucket_|tnf_kernel|pack_kernel|sub_apply' -s 30 -c 10 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
