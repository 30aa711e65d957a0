mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size" > gpurun_out/memcheck.log 2>&1; echo "exit $?"
grep -n "Invalid\|at 0x\|by thread\|=========     in \|Address\|kernel" gpurun_out/memcheck.log | head -40
tail -5 gpurun_out/memcheck.log
