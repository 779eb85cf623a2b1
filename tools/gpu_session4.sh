#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== dist check 2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist2.log 2>&1; echo "rc=$?"; grep -E "check|DIST_CHECK|Error|error" gpurun_out/dist2.log | tail -8
echo "== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "rc=$?"; tail -2 gpurun_out/bench_g2.json; tail -5 gpurun_out/bench_g2.err
echo "== bench 1 gpu"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; echo "rc=$?"; cat gpurun_out/bench_g1.json; tail -3 gpurun_out/bench_g1.err
