# round-2 GPU check: parity tests (not -x: see everything that fails), then the bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_last.log 2> gpurun_out/bench_last.err; echo "bench exit $?"
tail -3 gpurun_out/bench_last.err
python - <<'PY'
import json
l=open('gpurun_out/bench_last.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d.get('e2e',{}).get('value'), 'from_fastq', d.get('from_fastq'))
except Exception as e: print(l[-2000:])
PY
[ -n "$1" ] && timeout 600 $1 2>&1 | tee gpurun_out/extra.log
