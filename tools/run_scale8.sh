mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.log 2> gpurun_out/scale_$n.err
  python - <<PY
import json
l=open('gpurun_out/scale_$n.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l); print($n, d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d.get('e2e',{}).get('value'))
except Exception as e: print(l[-1500:]); print(open('gpurun_out/scale_$n.err').read()[-1500:])
PY
done
