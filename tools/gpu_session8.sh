#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== n18 on 8"; timeout 900 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 2 --warmup 1 --qubits 18 --layers 20 --no-e2e > gpurun_out/bench_n18.json 2> gpurun_out/bench_n18.err; echo "rc=$?"; tail -1 gpurun_out/bench_n18.json; grep -iE "error|Traceback" gpurun_out/bench_n18.err | tail -3
echo "== bench 8"; timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "rc=$?"; tail -1 gpurun_out/bench_g8.json; grep -iE "error|Traceback" gpurun_out/bench_g8.err | tail -3
echo "== bench 2"; timeout 900 $TR --nproc-per-node 2 --master-port 29526 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "rc=$?"; tail -1 gpurun_out/bench_g2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['passes_per_step'], d['nvlink'])"
