#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reserve1"; DMB_RESERVE_LOW=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "rc=$?"; cat gpurun_out/bench_r1.json
echo "== bench variant generic"; DMB_TILE_VARIANT=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo "rc=$?"; cat gpurun_out/bench_v1.json
echo "== probe"; timeout 400 python tools/gpu_probe.py > gpurun_out/probe2.jsonl 2> gpurun_out/probe2.err; echo "probe rc=$?"; tail -3 gpurun_out/probe2.err
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -o gpurun_out/prof_tile python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
