#!/bin/bash
# 1 GPU: parity of the compressed Sigma_c grid scan first, then the whole suite, timings of the three scan variants,
# ncu captures (TMA contraction kernel, compressed-scan kernels), launch list of a C60-shape step, C60 bench line.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gwbse.py -m gpu -x -q -k "grid" > gpurun_out/pytest_grid8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_grid8.log
tail -5 gpurun_out/pytest_grid8.log
timeout 300 python tools/bench_sigma_grid.py --workload synth-1000 --reps 2 --out gpurun_out/sigma_grid8.jsonl
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
tail -16 gpurun_out/pytest_gpu8.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c60_r8.json 2> gpurun_out/bench_c60_r8.err
tail -n 3 gpurun_out/bench_c60_r8.err; head -c 600 gpurun_out/bench_c60_r8.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_tma_kernel -c 2 -o gpurun_out/r01_contract_tma \
   python tools/bench_contract.py --reps 1 --nb 766 --naux 3830 --homo 72 --only epsilon_syrk,aux_rotation --out gpurun_out/sweep_ncu_tma.jsonl > gpurun_out/ncu_contract_tma.log 2>&1
XTPB_SIGMA_GRID=compressed timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:grid_compressed|ppm_moments' -c 2 -o gpurun_out/r01_sigma_grid_compressed \
   python tools/bench_sigma_grid.py --child --workload synth-500 --reps 1 > gpurun_out/ncu_grid_compressed.log 2>&1
XTPB_BENCH_MIN_WARMUP=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
   --log-file gpurun_out/launches_c60_r8.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_c60.log 2>&1
ls -la gpurun_out | tail -20
