#!/bin/bash
# Ring kernel (k_tile_ring6) parity + A/B on one B200:  gpurun --timeout 900 -- bash tools/gpu_ring_ab.sh
mkdir -p gpurun_out
T0=$SECONDS
VARS="${RING_VARIANTS:-2 3 4}"
OKV=""
for v in $VARS; do
  echo "== smoke variant $v"
  DMB_TILE_VARIANT=$v timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ring_smoke_$v.log 2>&1; rc=$?
  echo "rc=$rc t=$((SECONDS-T0))"; tail -2 gpurun_out/ring_smoke_$v.log
  if [ $rc -eq 0 ]; then OKV="$OKV $v"; fi
done
echo "variants alive:$OKV"
[ -z "$OKV" ] && exit 1
: > gpurun_out/r02_ring_ab.jsonl
for v in 0 $OKV 0 $OKV; do
  DMB_TILE_VARIANT=$v timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-side --no-parity 2> gpurun_out/ab.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({'variant': $v, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': d['roofline']['avg_launch_ms'], 'frac': d['roofline']['frac'], 'staging_only_ms': d['roofline'].get('staging_only', {}).get('ms'), 'one_gate_ms': d['roofline'].get('one_gate_per_launch', {}).get('ms'), 'prob_sum': d['prob_sum'], 'clocks': d['clocks']}))" >> gpurun_out/r02_ring_ab.jsonl
  tail -1 gpurun_out/r02_ring_ab.jsonl | cut -c1-400
done
echo "t=$((SECONDS-T0))"
echo "== parity of the ring variants"
DMB_TEST_TILE_VARIANT=$(echo $OKV | tr ' ' ',') timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tile_variant_parity" > gpurun_out/r02_ring_parity.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02_ring_parity.log
for v in $(echo $OKV | cut -d' ' -f1); do
  DMB_TILE_VARIANT=$v timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_tile_ring6 -s 20 -c 3 -f -o gpurun_out/r02_prof_ring_$v \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-side --no-parity > gpurun_out/ncu_ring_$v.log 2>&1; echo "ncu rc=$? t=$((SECONDS-T0))"
done
