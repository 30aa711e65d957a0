mkdir -p gpurun_out
run() { # policy batch_pairs tag
  PG_BIG_CACHE=$1 timeout 600 python bench.py --config c4 --batch-pairs $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_u_$3.log 2> gpurun_out/bench_u_$3.err; echo "$3 exit $?"
  python - $3 <<'PY'
import json,sys
f='gpurun_out/bench_u_%s.log'%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); st=d['roofline']['stages_ms']; print(sys.argv[1], d['value'], d['ms_per_step'], 'stage sum', round(sum(st.values()),1))
except Exception as e: print(f, 'failed', e)
PY
}
run lru 10000000 lru10
run lru 7000000 lru7
run largest 7000000 largest7
