timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^FAILED|^ERROR|passed|failed|^E  " | cut -c1-300 | head -20 > gpurun_out/r2k_pytest.log
cat gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-latency > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_bench.err
