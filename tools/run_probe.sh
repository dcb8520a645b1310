(cd tools && timeout 120 ./umma_shift64_probe > ../gpurun_out/r2e_shift64.log 2>&1)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^FAILED|^ERROR|passed|failed|^E  " | head -40 > gpurun_out/r2e_pytest.log
echo "== default" > gpurun_out/r2e_probe.log
RTOD_TC_TUNE_DBG=1 timeout 300 python tools/layer_probe.py L1 L2 L3 L5 L6 L7 L13 L99 >> gpurun_out/r2e_probe.log 2>&1
echo "== no split" >> gpurun_out/r2e_probe.log
RTOD_WSPLIT_AI=0 RTOD_TC_TUNE_DBG=1 timeout 300 python tools/layer_probe.py L1 L2 L3 L6 >> gpurun_out/r2e_probe.log 2>&1
cat gpurun_out/r2e_pytest.log
