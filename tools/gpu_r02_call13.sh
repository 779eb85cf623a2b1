#!/bin/bash
# Round 2, call 13 (1 GPU): direct global I/O of a pass's first / last op (dmb_io_op): parity, A/B by mask, ncu.
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02c_pytest_gpu.log
echo "== A/B direct I/O mask"
: > gpurun_out/r02c_direct_io_ab.jsonl
for mask in 3 0 1 2 3 0; do
  DMB_DIRECT_IO=$mask timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-side --no-parity 2> gpurun_out/ab.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); r = d['roofline']; print(json.dumps({'direct_io_mask': $mask, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': r['avg_launch_ms'], 'frac': r['frac'], 'smem_frac': r.get('shared_memory', {}).get('frac'), 'direct_io_ops_per_launch': r.get('direct_io_ops_per_launch'), 'staging_only_ms': r.get('staging_only', {}).get('ms'), 'one_gate_ms': r.get('one_gate_per_launch', {}).get('ms'), 'prob_sum': d['prob_sum'], 'clocks': d['clocks']}))" >> gpurun_out/r02c_direct_io_ab.jsonl
  tail -1 gpurun_out/r02c_direct_io_ab.jsonl; tail -2 gpurun_out/ab.err
done
echo "t=$((SECONDS-T0))"
echo "== bench (full line)"; timeout 600 python bench.py > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-400 gpurun_out/r02c_bench_1gpu.json; tail -3 gpurun_out/r02c_bench_1gpu.err
echo "== ncu full (3 launches of the tile kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass6 -s 20 -c 3 -f -o gpurun_out/r02c_prof_tile \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-side --no-parity > gpurun_out/ncu_full.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
ls -la gpurun_out | tail -8
