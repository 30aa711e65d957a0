# round 2, call F: scatter with the wait moved below the loads / index computation; auto choice of the featurize pass
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_f.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_f.log
tail -4 gpurun_out/pytest_f.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_f.log 2> gpurun_out/bench_f.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_f.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
except Exception as e: print('failed', e)
PY
