#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_ring_stagger.jsonl
for cfg in "2 0" "2 100" "2 250" "2 500" "2 1000" "3 250" "4 250" "0 0"; do
  set -- $cfg
  DMB_TILE_VARIANT=$1 DMB_RING_STAGGER_NS=$2 timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-side --no-parity 2> gpurun_out/ab.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': $1, 'stagger_ns': $2, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': d['roofline']['avg_launch_ms'], 'prob_sum': d['prob_sum']}))" >> gpurun_out/r02_ring_stagger.jsonl
  tail -1 gpurun_out/r02_ring_stagger.jsonl
done
