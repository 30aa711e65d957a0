# round 2, call C: tools tests, default bench (pread staging), ncu launch list + full capture of the count-side kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tools.py tests/test_gpu_ingest.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c.log
tail -25 gpurun_out/pytest_c.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.log 2> gpurun_out/bench_last.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_last.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'], d.get('from_fastq',{}).get('paths'))
except Exception as e: print('failed', e)
PY
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_r02.log 2>&1 && timeout 900 ncu --set full --import-source on --clock-control none -k regex:'bucket_s|sub_apply' -s 63 -c 4 -o gpurun_out/prof_r02_count -f $CMD > gpurun_out/ncu_r02_count.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
