#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu R3" ; DMB_TILE_VARIANT=4 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r3.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu_r3.log
for v in 0 4 5; do
echo "== bench variant $v"; DMB_TILE_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_v$v.err | tee gpurun_out/bench_v$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant',$v, d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['fused_ops_per_launch'], d['e2e']['ms_per_step'])"; tail -2 gpurun_out/bench_v$v.err
done
for mo in 6 8 12 16; do DMB_TILE_VARIANT=4 DMB_MAX_OPS_PER_PASS=$mo timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('r3 maxops',$mo, d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['fused_ops_per_launch'])"; done
echo "== ncu full R3"
DMB_TILE_VARIANT=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -o gpurun_out/prof_r3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_r3.log 2>&1; echo "ncu rc=$?"
