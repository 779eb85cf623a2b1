#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${NG:-2}
for plain in 0 1; do
echo "== bench $N plain=$plain"; DMB_EXCHANGE_PLAIN=$plain timeout 900 $TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e 2> gpurun_out/bench_p$plain.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('plain=$plain', d['value'], d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'])"; grep -iE "error|Traceback" gpurun_out/bench_p$plain.err | tail -3
done
