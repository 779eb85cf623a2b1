#!/bin/bash
mkdir -p gpurun_out
for v in 6 7 0; do
echo "== bench variant $v"; DMB_TILE_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_v$v.err | tee gpurun_out/bench_v$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant',$v, d['ms_per_step'], d['config']['passes_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms'])"; tail -2 gpurun_out/bench_v$v.err
done
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cat gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300
echo "== probe"; PROBE_VARIANTS=0 timeout 600 python tools/gpu_probe.py > gpurun_out/probe_final.jsonl 2> gpurun_out/probe_final.err; echo "probe rc=$?"
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
