#!/bin/bash
# 2-GPU check of the sharded path:  gpurun --gpus 2 --timeout 600 -- bash tools/gpu_2gpu.sh [bench-only]
mkdir -p gpurun_out
T0=$SECONDS
if [ "$1" != "bench-only" ]; then
echo "== pytest -m gpu (new tests + multi-GPU parity)"
timeout 300 python -m pytest tests/test_gpu_distributed.py tests/test_unitary.py tests/test_reference_quirks_switch.py tests/test_reduced_state.py -x -q -m gpu > gpurun_out/r02_pytest_gpu_2gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -5 gpurun_out/r02_pytest_gpu_2gpu.log
echo "== dist_check (2 ranks, vs oracle)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/harness/dist_check.py > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; grep -a "DIST_CHECK\|worst\|Error\|error\|sharded_dumps" gpurun_out/r02_dist_check_2gpu.log | tail -8
fi
echo "== bench --gpus 2 (reference arm, then ours)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_2gpu_reference.json 2> gpurun_out/bench_2gpu_ref.err; echo "rc=$? t=$((SECONDS-T0))"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$? t=$((SECONDS-T0))"; grep -a '"metric"' gpurun_out/r02_bench_2gpu.json | cut -c1-900; tail -5 gpurun_out/bench_2gpu.err
