#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/harness/dist_check.py > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "rc=$?"; grep -a "DIST_CHECK\|qft\|Error\|error\|pooled" gpurun_out/r02_dist_check_2gpu.log | tail -8 | cut -c1-400
