# round 2, call A: every GPU test (no -x), then the single-GPU runs of the other BASELINE configurations
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
for c in c3 c4; do
  timeout 1200 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/bench_r02_${c}_1gpu.json 2> gpurun_out/bench_r02_${c}_1gpu.err; echo "$c exit $?"
  tail -2 gpurun_out/bench_r02_${c}_1gpu.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_${c}_1gpu.json').read().strip().splitlines()[-1]); print('$c', d['value'], d['ms_per_step'], d['rows_total'], d['roofline']['stages_ms'], d['roofline']['whole_path'])
except Exception as e: print('$c failed', e)
PY
done
