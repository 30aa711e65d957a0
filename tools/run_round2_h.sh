# round 2, call H: second aggregation level in the sweep (A/B against the library without it), table depth 1..8, hash-mode record
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_h.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_h.log
tail -4 gpurun_out/pytest_h.log
for lib in agg noagg; do
  if [ $lib = noagg ]; then export PG_LIB_PATH=$PWD/pangaea_b200/libpangaea_b200_noagg.so; else unset PG_LIB_PATH; fi
  echo "== $lib" | tee -a gpurun_out/depth_r02.txt
  timeout 600 python tools/exp_table_depth.py 2>&1 | grep "table depth" | tee -a gpurun_out/depth_r02.txt
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_h_$lib.log 2> gpurun_out/bench_h_$lib.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_h_$lib.log').read().strip().splitlines()[-1]); print('$lib', d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
except Exception as e: print('failed', e)
PY
done
unset PG_LIB_PATH
timeout 900 python tools/exp_hash_mode.py > gpurun_out/hash_mode_r02.txt 2>&1; cat gpurun_out/hash_mode_r02.txt | tail -3
