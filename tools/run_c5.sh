mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --pairs 125000000 --steps 2 --warmup 3 --no-e2e > gpurun_out/c5_8gpu.log 2> gpurun_out/c5_8gpu.err
echo "exit $?"
python - <<PY
import json
l=open('gpurun_out/c5_8gpu.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['config'], d['roofline']['stages_ms'])
except Exception as e: print(l[-1500:]); print(open('gpurun_out/c5_8gpu.err').read()[-1500:])
PY
