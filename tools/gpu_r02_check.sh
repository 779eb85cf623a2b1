#!/bin/bash
# quick 1-GPU check of the committed library: GPU parity suite + one bench line
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02i_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02i_bench_1gpu.json 2> gpurun_out/r02i_bench_1gpu.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-300 gpurun_out/r02i_bench_1gpu.json; tail -3 gpurun_out/r02i_bench_1gpu.err
