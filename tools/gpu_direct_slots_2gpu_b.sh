#!/bin/bash
mkdir -p gpurun_out
echo "== dist_check, DMB_EXCHANGE_VARIANT=direct"
DMB_EXCHANGE_VARIANT=direct DIST_CHECK_FUZZ=12 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/harness/dist_check.py > gpurun_out/r02_dist_check_2gpu_direct.log 2>&1; echo "rc=$?"
grep -a "DIST_CHECK\|Error\|error" gpurun_out/r02_dist_check_2gpu_direct.log | tail -3 | cut -c1-300
for v in direct parked; do
  DMB_EXCHANGE_VARIANT=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 2 --no-side --no-e2e --no-cpu-baseline 2> gpurun_out/b2.err | grep -a '"metric"' | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': '$v', 'index': 'composed', 'ms_per_step': d['ms_per_step'], 'passes': d['config']['passes_per_step'], 'parity': d['parity_max_abs'], 'nvlink': d.get('nvlink')}))" | tee -a gpurun_out/r02_direct_slots_2gpu_b.jsonl
done
