#!/bin/bash
# 1 GPU: parity suite (BTDA, dipoles), ncu captures of the TMA contraction kernel and the Sigma_c grid kernel,
# launch list of a whole pentacene-shape step, final C60 bench line.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
tail -12 gpurun_out/pytest_gpu7.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_tma_kernel -c 2 -o gpurun_out/r01_contract_tma \
   python tools/bench_contract.py --reps 1 --nb 766 --naux 3830 --homo 72 --only epsilon_syrk,aux_rotation --out gpurun_out/sweep_ncu_tma.jsonl > gpurun_out/ncu_contract_tma.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sigma_ppm_grid_kernel -c 1 -o gpurun_out/r01_sigma_grid_v3 \
   python tools/bench_sigma_grid.py --child --workload synth-500 --reps 1 > gpurun_out/ncu_grid_v3.log 2>&1
XTPB_BENCH_MIN_WARMUP=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches_pentacene_final.csv python bench.py --workload pentacene-tzvp-shape --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c60_final.json 2> gpurun_out/bench_c60_final.err
tail -n 3 gpurun_out/bench_c60_final.err
