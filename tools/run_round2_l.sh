mkdir -p gpurun_out
PG_INGEST_DEBUG=1 timeout 1200 python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_l.log 2> gpurun_out/bench_l.err; echo "bench exit $?"
grep "pass 1\|pass 2" gpurun_out/bench_l.err | tail -6
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_l.log').read().strip().splitlines()[-1]); print(d['value'], (d.get('from_fastq') or {}).get('paths'))
except Exception as e: print('failed', e)
PY
