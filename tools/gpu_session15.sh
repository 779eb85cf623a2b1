#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench15.json')); print(d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['fused_ops_per_launch'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms'])"; tail -3 gpurun_out/bench15.err
for mo in 8 12 16; do DMB_MAX_OPS_PER_PASS=$mo timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maxops',$mo, d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['fused_ops_per_launch'])"; done
