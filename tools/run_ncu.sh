# usage: bash tools/run_ncu.sh <tag> [pairs] [skip] [count] - plain run first, then one ncu --set full capture of the hot kernels
# defaults capture one launch of every hot kernel of the 4th step at the headline workload (37 matching launches per step)
TAG=${1:-r01}; PAIRS=${2:-50000000}; SKIP=${3:-111}; COUNT=${4:-19}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --pairs $PAIRS --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'bucket_s|bucket_apply|tnf_kernel|pack_kernel|sub_apply' -s $SKIP -c $COUNT -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
