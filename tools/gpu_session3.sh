#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (subset)" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== probe"; timeout 600 python tools/gpu_probe.py > gpurun_out/probe3.jsonl 2> gpurun_out/probe3.err; echo "probe rc=$?"; tail -3 gpurun_out/probe3.err
for v in 0 2 3; do
echo "== bench variant $v"; DMB_TILE_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "rc=$?"; cat gpurun_out/bench_v$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
done
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -o gpurun_out/prof_tile3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
