mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -1 gpurun_out/bench_default.log | cut -c1-300
