mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gzip or paired or drop_in or golden" > gpurun_out/pytest_subset.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_subset.log
