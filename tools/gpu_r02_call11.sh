#!/bin/bash
# 1 GPU: the full -m gpu suite on the current tree + small-state timings (single-launch path on / off)
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02b_pytest_gpu.log
echo "== small configs, single-launch path on"; timeout 120 python tools/run_configs.py qft8 qft8_matrix grover12 > gpurun_out/r02b_configs_small_on.jsonl 2> gpurun_out/cfg.err; cat gpurun_out/r02b_configs_small_on.jsonl | cut -c1-700
echo "== small configs, single-launch path off"; DMB_SMALL_PATH=0 timeout 120 python tools/run_configs.py qft8 qft8_matrix > gpurun_out/r02b_configs_small_off.jsonl 2> gpurun_out/cfg.err; cat gpurun_out/r02b_configs_small_off.jsonl | cut -c1-700
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()"
echo "== sanitizer on the small path"
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02b_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02b_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r02b_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02b_sanitizer_racecheck.log
echo "t=$((SECONDS-T0))"
