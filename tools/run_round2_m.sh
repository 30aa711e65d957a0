mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_m.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_m.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_m.log 2> gpurun_out/bench_m.err; echo "bench exit $?"
tail -1 gpurun_out/bench_m.log | cut -c1-600
timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_m_c4.log 2> gpurun_out/bench_m_c4.err; echo "c4 exit $?"
tail -1 gpurun_out/bench_m_c4.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
