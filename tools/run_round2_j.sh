# round 2, call J: TNF overlap sweep, ingest staging rate, then the ncu launch list + captures of the final kernels
mkdir -p gpurun_out
for n in 1 2 3; do
  PG_TNF_OVERLAP=$n timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_j_tnf$n.log 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_j_tnf$n.log').read().strip().splitlines()[-1]); print('tnf_overlap', $n, d['value'], d['ms_per_step'], d['roofline']['stages_ms']['tnf'], d['roofline']['stages_ms']['feat_apply'])
except Exception as e: print('failed', e)
PY
done
PG_INGEST_DEBUG=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_j_ingest.log 2> gpurun_out/bench_j_ingest.err
grep "staged" gpurun_out/bench_j_ingest.err | tail -6
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_j_ingest.log').read().strip().splitlines()[-1]); print(d['value'], (d.get('from_fastq') or {}).get('paths'))
except Exception as e: print('failed', e)
PY
bash tools/run_ncu.sh r02
