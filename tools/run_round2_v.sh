mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_v.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v.log 2> gpurun_out/bench_v.err; echo "bench exit $?"
timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_v_c4.log 2> gpurun_out/bench_v_c4.err; echo "c4 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_v.log','gpurun_out/bench_v_c4.log'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
    except Exception as e: print(f, 'failed', e)
PY
