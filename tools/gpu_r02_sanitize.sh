#!/bin/bash
# compute-sanitizer on the final library (chained kernels, cluster path, replay)
mkdir -p gpurun_out
T0=$SECONDS
timeout 280 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02i_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$? t=$((SECONDS-T0))"; grep -a "sanitize_smoke\|ERROR SUMMARY" gpurun_out/r02i_sanitizer_memcheck.log | tail -8
timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r02i_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$? t=$((SECONDS-T0))"; grep -a "RACECHECK SUMMARY\|hazard" gpurun_out/r02i_sanitizer_racecheck.log | tail -4
