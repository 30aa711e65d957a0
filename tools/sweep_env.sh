# usage: bash tools/sweep_env.sh VAR v1 v2 ...
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_env.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/sweep_env.log').read().strip().splitlines()[-1]); s=d['roofline']['stages_ms']; print("$VAR=$v", round(d['value']/1e6,1), d['ms_per_step'], 'tnf', s['tnf'], 'feat_apply', s['feat_apply'])
PY
done
