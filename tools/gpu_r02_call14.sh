#!/bin/bash
# Round 2, call 14 (1 GPU): direct I/O A/B with the spill-free 4-CTA instantiations (DMB_DIO_CTAS=4) against 5 CTAs.
mkdir -p gpurun_out
T0=$SECONDS
: > gpurun_out/r02c_direct_io_ab2.jsonl
for cfg in "0 5" "3 4" "1 4" "2 4" "3 5" "3 4" "0 5"; do
  set -- $cfg
  DMB_DIRECT_IO=$1 DMB_DIO_CTAS=$2 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-side --no-parity 2> gpurun_out/ab.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); r = d['roofline']; print(json.dumps({'direct_io_mask': $1, 'dio_ctas': $2, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': r['avg_launch_ms'], 'frac': r['frac'], 'smem_frac': r.get('shared_memory', {}).get('frac'), 'direct_io_ops_per_launch': r.get('direct_io_ops_per_launch'), 'staging_only_ms': r.get('staging_only', {}).get('ms'), 'one_gate_ms': r.get('one_gate_per_launch', {}).get('ms'), 'prob_sum': d['prob_sum'], 'parity': d.get('parity_max_abs'), 'clocks': d['clocks']}))" >> gpurun_out/r02c_direct_io_ab2.jsonl
  tail -1 gpurun_out/r02c_direct_io_ab2.jsonl; tail -2 gpurun_out/ab.err
done
echo "t=$((SECONDS-T0))"
