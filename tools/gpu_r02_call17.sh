#!/bin/bash
# Round 2, call 17 (1 GPU): chained ops (two CNOTs of a cu1 in one round trip): parity, QFT-16 A/B, config 3 unchanged.
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02e_pytest_gpu.log
echo "== A/B chained ops"
: > gpurun_out/r02e_chain_ab.jsonl
for cfg in "qft16 1" "qft16 0" "config3 1" "qft16 1" "qft16 0" "config3 0"; do
  set -- $cfg
  DMB_CHAIN_OPS=$2 timeout 200 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-side --no-parity 2> gpurun_out/ab.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read()); r = d['roofline']; print(json.dumps({'workload': '$1', 'chain_ops': $2, 'ms_per_step': d['ms_per_step'], 'avg_launch_ms': r['avg_launch_ms'], 'frac': r['frac'], 'smem_frac': r.get('shared_memory', {}).get('frac'), 'fused_ops_per_launch': r['fused_ops_per_launch'], 'chained_per_launch': r.get('ops_chained_to_their_predecessor_per_launch'), 'prob_sum': d['prob_sum'], 'clocks': d['clocks']}))" >> gpurun_out/r02e_chain_ab.jsonl
  tail -1 gpurun_out/r02e_chain_ab.jsonl; tail -2 gpurun_out/ab.err
done
echo "t=$((SECONDS-T0))"
echo "== small configs"; timeout 200 python tools/run_configs.py qft8 grover12 > gpurun_out/r02e_configs_small.jsonl 2> gpurun_out/cfg.err; cut -c1-260 gpurun_out/r02e_configs_small.jsonl; tail -3 gpurun_out/cfg.err
