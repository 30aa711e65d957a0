# round 2, call E: scatter v3 (reservations; 256- vs 512-word tiles) x featurize path (lookup+collect vs round-1 sweep)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_e.log
tail -6 gpurun_out/pytest_e.log
for lib in t256 t512; do for mode in 0 1; do
  if [ $lib = t512 ]; then export PG_LIB_PATH=$PWD/pangaea_b200/libpangaea_b200_t512.so; else unset PG_LIB_PATH; fi
  export PG_FEAT_APPLY=$mode
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_e_${lib}_$mode.log 2> gpurun_out/bench_e_${lib}_$mode.err; echo "bench $lib $mode exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_e_${lib}_$mode.log').read().strip().splitlines()[-1]); print('$lib', '$mode', d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
except Exception as e: print('failed', e)
PY
done; done
