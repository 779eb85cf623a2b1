#!/bin/bash
# Round 2, final 1-GPU session: what the driver runs (pytest -m gpu, smoke, both bench arms) + the ncu records of the final library.
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/r02f_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench --impl reference"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2> gpurun_out/ref.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-300 gpurun_out/r02f_bench_reference_arm.json
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err; echo "rc=$? t=$((SECONDS-T0))"; cut -c1-500 gpurun_out/r02f_bench_1gpu.json; tail -3 gpurun_out/r02f_bench_1gpu.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
echo "== ncu full (3 launches of the tile kernel, config 3)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass6 -s 20 -c 3 -f -o gpurun_out/r02f_prof_tile \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-side --no-parity > gpurun_out/ncu_full.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
echo "== ncu full (3 launches of the chained tile kernel, QFT-16)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass6 -s 6 -c 3 -f -o gpurun_out/r02f_prof_tile_qft16 \
    python bench.py --workload qft16 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-side --no-parity > gpurun_out/ncu_full_qft.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"
ls -la gpurun_out | tail -6
