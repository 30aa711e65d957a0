mkdir -p gpurun_out
for v in "1 2" "0 2" "1 0" "0 0" "1 1"; do
  set -- $v
  PG_TNF_FOLD=$1 PG_TNF_OVERLAP=$2 timeout 600 python bench.py --config c4 --pairs 100000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_p_$1_$2.log 2> gpurun_out/bench_p_$1_$2.err; echo "fold=$1 overlap=$2 exit $?"
  python - "$1" "$2" <<'PY'
import json,sys
f='gpurun_out/bench_p_%s_%s.log'%(sys.argv[1],sys.argv[2])
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
except Exception as e: print(f, 'failed', e)
PY
done
