#!/bin/bash
# 2-GPU check of the sharded path:  gpurun --gpus 2 --timeout 300 -- bash tools/gpu_2gpu.sh [bench-only]
mkdir -p gpurun_out
T0=$SECONDS
if [ "$1" != "bench-only" ]; then
echo "== dist_check (2 ranks, vs oracle)"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/harness/dist_check.py > gpurun_out/dist_check_2gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; grep -a "DIST_CHECK\|worst\|Error\|error" gpurun_out/dist_check_2gpu.log | tail -5
fi
echo "== bench --gpus 2"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$? t=$((SECONDS-T0))"; grep -a '"metric"' gpurun_out/bench_2gpu.json | cut -c1-700
