#!/bin/bash
# Round 2, call 15 (1 GPU): TMA staging probe (tools/probes/tma_probe.cu), small configs with the compiled replay.
mkdir -p gpurun_out
T0=$SECONDS
echo "== tma probe"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe tools/probes/tma_probe.cu && timeout 300 /tmp/tma_probe | tee gpurun_out/r02_tma_probe.jsonl
echo "rc=$? t=$((SECONDS-T0))"
echo "== small configs"; timeout 200 python tools/run_configs.py qft8 grover12 > gpurun_out/r02c_configs_small.jsonl 2> gpurun_out/cfg.err; cut -c1-330 gpurun_out/r02c_configs_small.jsonl; tail -3 gpurun_out/cfg.err
echo "t=$((SECONDS-T0))"
