for lib in "$@"; do
  PG_LIB_PATH=$PWD/$lib timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_lib.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/sweep_lib.log').read().strip().splitlines()[-1]); print("$lib", d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
PY
done
