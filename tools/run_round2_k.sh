# round 2, call K: pipelined device ingest: ingest tests + default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k.log
tail -5 gpurun_out/pytest_k.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_last.log 2> gpurun_out/bench_last.err; echo "bench exit $?"
tail -2 gpurun_out/bench_last.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_last.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), (d.get('from_fastq') or {}), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('failed', e)
PY
