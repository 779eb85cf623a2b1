#!/bin/bash
# Round 2, call 2 (1 GPU): parity on the new default kernel, the bench line, CTA-count / fold A/B, DMMA probe, ncu.
#   gpurun --timeout 1500 -- bash tools/gpu_r02_call2.sh
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; DMB_TEST_TILE_VARIANT=2,3 timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02_pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-600 gpurun_out/r02_bench_1gpu.json; tail -5 gpurun_out/r02_bench_1gpu.err
echo "== A/B"
: > gpurun_out/r02_tile_variants.jsonl
for cfg in "0 1" "2 1" "3 1" "0 0" "1 1" "0 1" "3 1"; do
  set -- $cfg
  DMB_TILE_VARIANT=$1 DMB_FOLD_TSP0=$2 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-side --no-parity 2> gpurun_out/ab.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': $1, 'fold_tsp0': $2, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': d['roofline']['avg_launch_ms'], 'frac': d['roofline']['frac'], 'smem_frac': d['roofline'].get('shared_memory', {}).get('frac'), 'staging_only_ms': d['roofline'].get('staging_only', {}).get('ms'), 'one_gate_ms': d['roofline'].get('one_gate_per_launch', {}).get('ms'), 'clocks': d['clocks']}))" >> gpurun_out/r02_tile_variants.jsonl
  tail -1 gpurun_out/r02_tile_variants.jsonl
done
echo "t=$((SECONDS-T0))"
echo "== dmma probe"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/probes/dmma_probe.cu && /tmp/dmma_probe | tee gpurun_out/r02_dmma_probe.jsonl
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
echo "== configs"; timeout 300 python tools/run_configs.py qft8 grover12 > gpurun_out/r02_configs_1gpu.jsonl 2> gpurun_out/configs.err; echo "rc=$?"; cat gpurun_out/r02_configs_1gpu.jsonl | cut -c1-300
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
echo "== ncu full (3 launches of the tile kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass6 -s 20 -c 3 -f -o gpurun_out/r02_prof_tile \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-side --no-parity > gpurun_out/ncu_full.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
ls -la gpurun_out | tail -12
