for sw in 67108864 134217728; do
  PG_FEAT_SEG_WORDS=$sw timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$sw.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/sweep_$sw.log').read().strip().splitlines()[-1]); print($sw, d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
