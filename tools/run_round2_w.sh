mkdir -p gpurun_out
PG_BENCH_DEBUG=1 timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_w_c4.log 2> gpurun_out/bench_w_c4.err; echo "c4 exit $?"
grep "\[bench\]" gpurun_out/bench_w_c4.err | tail -90 | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w_c4.log').read().strip().splitlines()[-1]); st=d['roofline']['stages_ms']; print(d['value'], d['ms_per_step'], 'stage sum', round(sum(st.values()),1))
PY
