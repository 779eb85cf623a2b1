#!/bin/bash
# 8-GPU box: multi-rank pytest, scaling bench at 2/4/8 (fused NVLink exchange), NCCL A/B at 8, config 5 (n=18), QFT-16.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== pytest gpu distributed"; timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/pytest_gpu_dist.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_dist.log
for N in 8 4 2; do
echo "== bench $N"; timeout 900 $TR --nproc-per-node $N --master-port $((29530+N)) bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_final_g$N.json 2> gpurun_out/bench_final_g$N.err; echo "rc=$?"; tail -1 gpurun_out/bench_final_g$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($N, d['value'], d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e'] and d['e2e']['ms_per_step'], d['nvlink'])"; grep -iE "error|Traceback" gpurun_out/bench_final_g$N.err | tail -3
done
echo "== bench 8 nccl exchange"; DMB_FUSED_EXCHANGE=0 timeout 900 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nccl8', d['value'], d['ms_per_step'])"
echo "== n18 on 8"; timeout 900 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 2 --warmup 1 --qubits 18 --layers 20 --no-e2e > gpurun_out/bench_final_n18.json 2> gpurun_out/bench_final_n18.err; echo "rc=$?"; tail -1 gpurun_out/bench_final_n18.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n18', d['value'], d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['nvlink'])"; grep -iE "error|Traceback" gpurun_out/bench_final_n18.err | tail -3
for N in 8 2; do
echo "== qft16 on $N"; timeout 600 $TR --nproc-per-node $N --master-port $((29550+N)) tools/run_configs.py qft16 > gpurun_out/configs_qft16_g$N.jsonl 2> gpurun_out/configs_qft16_g$N.err; echo "rc=$?"; cat gpurun_out/configs_qft16_g$N.jsonl | cut -c1-500
done
