# usage: bash tools/run_ncu.sh <tag>
#  1. plain run (must exit 0)
#  2. launch list: every launch of the timed steps with its device time (cold-cache, serialised: compare SHARES)
#  3. ncu --set full of one launch of every hot kernel of the 4th step at the headline workload (after 3 warm-up steps).
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_${TAG}_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'bucket_s|pack_kernel|sub_apply' -s 63 -c 5 -o gpurun_out/prof_${TAG}_count -f $CMD > gpurun_out/ncu_${TAG}_count.log 2>&1
echo "exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'tnf_kernel|bucket_apply_feat|normalize_rows' -s 21 -c 4 -o gpurun_out/prof_${TAG}_feat -f $CMD > gpurun_out/ncu_${TAG}_feat.log 2>&1
echo "exit $?"; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$TAG.csv
