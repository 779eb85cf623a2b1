#!/bin/bash
# A/B of the tile-kernel variants under the tile-search schedule (~11.5 fused ops per launch):
#   gpurun --timeout 400 -- bash tools/gpu_variants.sh
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for v in 0 6 7 2 3; do
  DMB_TILE_VARIANT=$v timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/variant_$v.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': $v, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': d['roofline']['avg_launch_ms'], 'passes': d['config']['passes_per_step'], 'clocks': d['clocks']}))" >> gpurun_out/variants.jsonl
  tail -1 gpurun_out/variants.jsonl
done
