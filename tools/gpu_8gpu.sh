#!/bin/bash
# 8-GPU validation (charged 8x): sharded parity vs the oracle, then both bench arms as the driver launches them.
#   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_8gpu.sh
mkdir -p gpurun_out
T0=$SECONDS
N=${NGPU:-8}
echo "== dist_check ($N ranks, vs oracle)"
DIST_CHECK_FUZZ=${DIST_CHECK_FUZZ:-16} timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tests/harness/dist_check.py > gpurun_out/r02_dist_check_${N}gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; grep -a "DIST_CHECK\|Error\|error\|sharded_dumps\|random_programs" gpurun_out/r02_dist_check_${N}gpu.log | tail -6 | cut -c1-300
echo "== bench --impl reference --gpus $N"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_${N}gpu_reference.json 2> gpurun_out/bench_ref.err; echo "rc=$? t=$((SECONDS-T0))"
echo "== bench --gpus $N"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$? t=$((SECONDS-T0))"; grep -a '"metric"' gpurun_out/r02_bench_${N}gpu.json | cut -c1-1500; tail -5 gpurun_out/bench_${N}gpu.err | cut -c1-300
