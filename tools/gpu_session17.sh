#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['staging_only'], d['roofline']['one_gate_per_launch'], d['e2e'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench_final.err
echo "== probe"; PROBE_VARIANTS=0 timeout 600 python tools/gpu_probe.py > gpurun_out/probe_final.jsonl 2> gpurun_out/probe_final.err; echo "probe rc=$?"
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
