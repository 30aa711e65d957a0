# round 2, call B: new tests (sampler, compact), c3 on one GPU, bench default (from_fastq with both ingest paths)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sampler.py tests/test_gpu_ingest.py tests/test_gpu_parity.py -m gpu -q -x -k "sampler or device or compacted or batches" > gpurun_out/pytest_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_b.log
tail -15 gpurun_out/pytest_b.log
timeout 1200 python bench.py --config c3 --steps 2 --warmup 3 > gpurun_out/bench_r02_c3_1gpu.json 2> gpurun_out/bench_r02_c3_1gpu.err; echo "c3 exit $?"
tail -2 gpurun_out/bench_r02_c3_1gpu.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_last.log 2> gpurun_out/bench_last.err; echo "bench exit $?"
tail -3 gpurun_out/bench_last.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_r02_c3_1gpu.json','gpurun_out/bench_last.log'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['stages_ms'], d.get('from_fastq'))
    except Exception as e: print(f, 'failed', e)
PY
