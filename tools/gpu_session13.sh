#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${NG:-2}
echo "== dist check push"; DMB_EXCHANGE=push timeout 600 $TR --nproc-per-node $N --master-port 29511 tools/dist_check.py > gpurun_out/dist_push.log 2>&1; echo "rc=$?"; grep -E "DIST_CHECK|Error|error" gpurun_out/dist_push.log | tail -4
for mode in pull push; do
echo "== bench $N $mode"; DMB_EXCHANGE=$mode timeout 900 $TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e 2> gpurun_out/bench_x_$mode.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', d['value'], d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'])"; grep -iE "error|Traceback" gpurun_out/bench_x_$mode.err | tail -3
done
if [ "$N" = "8" ]; then
echo "== n18 push"; DMB_EXCHANGE=push timeout 900 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 2 --warmup 1 --qubits 18 --layers 20 --no-e2e 2> gpurun_out/bench_n18_push.err | tail -1 | tee gpurun_out/bench_n18_push.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n18 push', d['value'], d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'])"
fi
