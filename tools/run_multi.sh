N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_multi.log
for n in 1 $N; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.log 2>&1
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.log 2>&1; fi
  echo "bench N=$n exit $?"
  python - <<PY
import json
l=open('gpurun_out/scale_$n.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l); print($n, d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d.get('e2e',{}).get('value'))
except Exception as e: print(l[-1500:])
PY
done
