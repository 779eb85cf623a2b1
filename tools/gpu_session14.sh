#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== bench full"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cat gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
