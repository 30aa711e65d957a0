mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
l=open('gpurun_out/bench_last.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d.get('e2e',{}).get('value'))
except Exception as e: print(l[-2000:])
PY
[ -n "$1" ] && timeout 300 $1 2>&1 | tee gpurun_out/extra.log
