mkdir -p gpurun_out
timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_t_c4.log 2> gpurun_out/bench_t_c4.err; echo "c4 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_t_c4.log',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); st=d['roofline']['stages_ms']; print(d['value'], d['ms_per_step'], 'stage sum', round(sum(v for k,v in st.items()),1), st)
    except Exception as e: print(f, 'failed', e)
PY
