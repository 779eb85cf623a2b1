#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02j_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
