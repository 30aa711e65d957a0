# usage: bash tools/run_multi.sh N "c2 c3"   - multi-GPU bench lines (and, at N = 2, the NCCL parity tests)
N=${1:-2}; CFGS=${2:-c2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
  tail -6 gpurun_out/pytest_multi.log
fi
for c in $CFGS; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 3 --warmup 3 \
     > gpurun_out/bench_r02_${c}_${N}gpu.json 2> gpurun_out/bench_r02_${c}_${N}gpu.err; echo "$c x$N exit $?"
  tail -3 gpurun_out/bench_r02_${c}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_${c}_${N}gpu.json').read().strip().splitlines()[-1]); print('$c', $N, d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', (d.get('e2e') or {}).get('value'), 'parity', d.get('parity_check'))
except Exception as e: print('$c failed', e)
PY
done
