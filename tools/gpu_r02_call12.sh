#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02b_pytest_gpu.log
echo "== small configs"; timeout 120 python tools/run_configs.py qft8 grover12 > gpurun_out/r02b_configs_small_on.jsonl 2> gpurun_out/cfg.err; cat gpurun_out/r02b_configs_small_on.jsonl | cut -c1-400
echo "== small configs, single-launch path off"; DMB_SMALL_PATH=0 timeout 120 python tools/run_configs.py qft8 > gpurun_out/r02b_configs_small_off.jsonl 2> gpurun_out/cfg.err; cat gpurun_out/r02b_configs_small_off.jsonl | cut -c1-400
echo "t=$((SECONDS-T0))"
