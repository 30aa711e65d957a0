# round 2, call D: lookup + collect featurize: parity tests, then bench (default, A/B with the old sweep), c4
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_d.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_d.log
tail -15 gpurun_out/pytest_d.log
for mode in new old; do
  if [ $mode = old ]; then export PG_FEAT_APPLY=1; else unset PG_FEAT_APPLY; fi
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_d_$mode.log 2> gpurun_out/bench_d_$mode.err; echo "bench $mode exit $?"
  tail -2 gpurun_out/bench_d_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_d_$mode.log').read().strip().splitlines()[-1]); print('$mode', d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
except Exception as e: print('failed', e)
PY
done
unset PG_FEAT_APPLY
timeout 1200 python bench.py --config c4 --steps 1 --warmup 3 > gpurun_out/bench_r02_c4_1gpu.json 2> gpurun_out/bench_r02_c4_1gpu.err; echo "c4 exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_c4_1gpu.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'], d['roofline']['stages_ms'], d['checksum'])
except Exception as e: print('failed', e)
PY
