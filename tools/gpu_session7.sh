#!/bin/bash
# 8-GPU session: correctness of the sharded engine at world 8, scaling bench, BASELINE config 5 (n=18).
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== dist check 8"; timeout 600 $TR --nproc-per-node 8 --master-port 29521 tools/dist_check.py > gpurun_out/dist8.log 2>&1; echo "rc=$?"; grep -E "check|DIST_CHECK|Error" gpurun_out/dist8.log | tail -6
echo "== dist check 4"; timeout 600 $TR --nproc-per-node 4 --master-port 29522 tools/dist_check.py > gpurun_out/dist4.log 2>&1; echo "rc=$?"; grep -E "DIST_CHECK|Error" gpurun_out/dist4.log | tail -3
echo "== bench 8"; timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "rc=$?"; tail -1 gpurun_out/bench_g8.json; grep -iE "error|Traceback" gpurun_out/bench_g8.err | tail -3
echo "== bench 4"; timeout 900 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "rc=$?"; tail -1 gpurun_out/bench_g4.json; grep -iE "error|Traceback" gpurun_out/bench_g4.err | tail -3
echo "== n18 on 8"; timeout 900 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 2 --warmup 1 --n 18 --depth 20 --no-e2e > gpurun_out/bench_n18.json 2> gpurun_out/bench_n18.err; echo "rc=$?"; tail -1 gpurun_out/bench_n18.json; grep -iE "error|Traceback" gpurun_out/bench_n18.err | tail -3
