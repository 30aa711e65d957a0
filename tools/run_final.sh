# final single-GPU check of a round: GPU tests, the driver's default bench line, the streamed configurations, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench exit $?"
timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_final_c4.log 2> gpurun_out/bench_final_c4.err; echo "c4 exit $?"
timeout 600 python bench.py --config c3 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_final_c3.log 2> gpurun_out/bench_final_c3.err; echo "c3 exit $?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_ref.log 2> gpurun_out/bench_final_ref.err; echo "reference arm exit $?"
python - <<'PY'
import json
for f in ('bench_final','bench_final_c4','bench_final_c3','bench_final_ref'):
    try:
        d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d.get('ms_per_step'), (d.get('roofline') or {}).get('stages_ms'), (d.get('e2e') or {}).get('value'), ((d.get('from_fastq') or {}).get('paths') or {}).get('device_ingest'))
    except Exception as e: print(f, 'failed', e)
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
