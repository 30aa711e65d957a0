for i in 1 2 3 4; do
  timeout 600 python bench.py --steps ${1:-3} --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/rep_$i.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/rep_$i.log').read().strip().splitlines()[-1]); print($i, d['value'], d['ms_per_step'], d['device_ms_per_step'], sum(d['roofline']['stages_ms'].values()), d['roofline']['stages_ms']['feat_apply'])
PY
done 2>/dev/null
