#!/bin/bash
# One GPU session that regenerates everything under profiles/ that comes from a single B200:
#   gpurun --timeout 1500 -- bash tools/gpu_session.sh
# then, back in the build container:  python tools/summarize_profiles.py
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_final.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
echo "== probes"; timeout 300 python tools/gpu_probe.py > gpurun_out/probe_final.jsonl 2> gpurun_out/probe_final.err; echo "rc=$?"
timeout 200 python tools/store_probe.py > gpurun_out/store_probe.jsonl 2>&1
echo "== configs"; timeout 600 python tools/run_configs.py qft8 qft8_matrix grover12 layered14 qft16 > gpurun_out/configs_final.jsonl 2> gpurun_out/configs_final.err; echo "rc=$?"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full (3 launches of the tile kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -f -o gpurun_out/prof_final \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -15
