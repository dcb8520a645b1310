echo "== tc kernel heuristic config, trace stamps" > gpurun_out/r2h_probe.log
for L in L13 L38 L63 L64; do
RTOD_TC_NO_PAIR=1 PROBE_FLAGS=4 RTOD_LIB=$PWD/realtimeobjectdetection_b200/librtod_trace.so RTOD_CLK_DBG=1 PROBE_REPS=2 timeout 120 python tools/layer_probe.py $L 2>&1 | tail -6 >> gpurun_out/r2h_probe.log
done
echo "== same, no trace" >> gpurun_out/r2h_probe.log
RTOD_TC_NO_PAIR=1 PROBE_FLAGS=4 timeout 120 python tools/layer_probe.py L13 L38 L63 L64 >> gpurun_out/r2h_probe.log 2>&1
