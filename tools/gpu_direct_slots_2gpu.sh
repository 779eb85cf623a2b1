#!/bin/bash
# 2 GPUs: the sharded parity harness with each exchange-schedule variant forced (small states never pick them on cost)
mkdir -p gpurun_out
for v in direct direct_late; do
  echo "== dist_check, DMB_EXCHANGE_VARIANT=$v"
  DMB_EXCHANGE_VARIANT=$v DIST_CHECK_FUZZ=24 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/harness/dist_check.py > gpurun_out/r02_dist_check_2gpu_$v.log 2>&1; echo "rc=$?"
  grep -a "DIST_CHECK\|Error\|error\|pooled\|random_programs" gpurun_out/r02_dist_check_2gpu_$v.log | tail -5 | cut -c1-300
done
echo "== bench --gpus 2 --workload layered16, variants"
for v in parked direct direct_late; do
  DMB_EXCHANGE_VARIANT=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 2 --no-side --no-e2e --no-cpu-baseline 2> gpurun_out/b2.err | grep -a '"metric"' | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': '$v', 'ms_per_step': d['ms_per_step'], 'passes': d['config']['passes_per_step'], 'parity': d['parity_max_abs'], 'nvlink': d.get('nvlink')}))" | tee -a gpurun_out/r02_direct_slots_2gpu.jsonl
done
