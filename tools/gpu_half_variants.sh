#!/bin/bash
# Round-2 first call: parity + A/B of the half-CTA tile kernel (dmb_set_tile_variant 8 / 9):
#   gpurun --timeout 1100 -- bash tools/gpu_half_variants.sh
# 1. GPU parity tests with the variant forced (golden cases, random programs, n = 14 properties)
# 2. bench A/B in one session: default (0) vs 8 / 9 (half-CTA, 6 / 5 CTAs x 128 threads) vs 10 / 11 / 12 (paired bodies: 4 / 5 CTAs x 1 stage, 3 CTAs x 2 stages; 13 = 10 + TSP factor folded into the control map)
mkdir -p gpurun_out
T0=$SECONDS
for v in 8 10 12 13 14; do
  echo "== parity, variant $v"
  DMB_TEST_TILE_VARIANT=$v timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "experimental_tile_variant" > gpurun_out/half_parity_$v.log 2>&1
  echo "rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/half_parity_$v.log
done
: > gpurun_out/half_variants.jsonl
for v in 0 14 8 9 10 11 12 13 0 14 8 10 13; do
  DMB_TILE_VARIANT=$v timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/half_variant_$v.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': $v, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': d['roofline']['avg_launch_ms'], 'passes': d['config']['passes_per_step'], 'clocks': d['clocks']}))" >> gpurun_out/half_variants.jsonl
  tail -1 gpurun_out/half_variants.jsonl
done
echo "t=$((SECONDS-T0))"
# 3. ncu --set full of two launches of the half-CTA and the paired kernel (stall reasons, pipe utilisation, registers)
for v in 8 10; do
  DMB_TILE_VARIANT=$v timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass6_half -s 20 -c 2 -f \
      -o gpurun_out/prof_variant_$v python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_variant_$v.log 2>&1
  echo "ncu variant $v rc=$? t=$((SECONDS-T0))"
done
ls -la gpurun_out | tail -15
