# round 2, call I: TNF whole-cloud stores, stash pool: all GPU tests; default bench; c4; ingest phase times; hash-mode record
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_i.log
tail -6 gpurun_out/pytest_i.log
PG_INGEST_DEBUG=1 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_last.log 2> gpurun_out/bench_last.err; echo "bench exit $?"
grep "\[ingest\]" gpurun_out/bench_last.err | tail -24
timeout 1200 python bench.py --config c4 --steps 1 --warmup 3 > gpurun_out/bench_r02_c4_1gpu.json 2> gpurun_out/bench_r02_c4_1gpu.err; echo "c4 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_last.log','gpurun_out/bench_r02_c4_1gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', (d.get('e2e') or {}).get('value'), (d.get('from_fastq') or {}).get('paths'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'failed', e)
PY
timeout 900 python tools/exp_hash_mode.py > gpurun_out/hash_mode_r02.txt 2>&1; tail -3 gpurun_out/hash_mode_r02.txt
