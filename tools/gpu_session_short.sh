#!/bin/bash
# Short single-GPU session (fits ~10 box minutes):  gpurun --timeout 780 -- bash tools/gpu_session_short.sh
# bench (new scheduler), A/B bench with the previous schedule, GPU parity tests, small-config wall
# times, ncu launch list, ncu --set full of three tile-pass launches.
mkdir -p gpurun_out
T0=$SECONDS
echo "== bench"; timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-600 gpurun_out/bench_final.json
echo "== bench, previous schedule (program order, cap 10)"
DMB_SCHED_STRATEGY=0 DMB_MAX_OPS_PER_PASS=10 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_prev_schedule.json 2> gpurun_out/bench_prev_schedule.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-300 gpurun_out/bench_prev_schedule.json
echo "== pytest gpu"; timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/pytest_gpu.log
echo "== ncu launch list"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
echo "== configs"; timeout 150 python tools/run_configs.py qft8 qft8_matrix grover12 > gpurun_out/configs_final.jsonl 2> gpurun_out/configs_final.err; echo "rc=$? t=$((SECONDS-T0))"; cat gpurun_out/configs_final.jsonl | cut -c1-300
echo "== ncu full (3 launches of the tile kernel)"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -f -o gpurun_out/prof_final \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
ls -la gpurun_out | tail -12
