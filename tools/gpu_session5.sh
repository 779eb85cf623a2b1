#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "rc=$?"; cat gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
echo "== bench maxops variants"
for mo in 3 4 5 6; do DMB_MAX_OPS_PER_PASS=$mo timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maxops',$mo, d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms'])"; done
echo "== reserve1"; DMB_RESERVE_LOW=1 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('reserve1', d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
echo "== probe"; PROBE_VARIANTS=0 timeout 600 python tools/gpu_probe.py > gpurun_out/probe5.jsonl 2> gpurun_out/probe5.err; echo "probe rc=$?"; tail -3 gpurun_out/probe5.err
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -o gpurun_out/prof_tile5 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
