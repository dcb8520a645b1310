#!/bin/bash
# Runs every bring-up stage in its own process with a time limit; logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for stage in "$@"; do
  timeout 300 python tools/gpu_probe.py "$stage" > gpurun_out/probe_$stage.log 2>&1
  echo "stage $stage exit $?" | tee -a gpurun_out/probe_$stage.log
  tail -n 60 gpurun_out/probe_$stage.log
done
