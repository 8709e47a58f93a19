#!/bin/bash
# 1 GPU: parity suite, Sigma_c grid-kernel variants, contraction sweep, ncu captures of both hot kernels, C60 bench.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -5 gpurun_out/pytest_gpu3.log
timeout 600 python tools/bench_sigma_grid.py --workload synth-1000 --out gpurun_out/sigma_grid.jsonl
timeout 600 python tools/bench_contract.py --reps 5 --out gpurun_out/contract_sweep_v1.jsonl > gpurun_out/sweep_v1.log 2>&1
for g in 8 1; do
XTPB_GRID_GROUP=$g timeout 600 ncu --set full --clock-control none --import-source on -k regex:sigma_ppm_grid -c 1 -o gpurun_out/r01_sigma_grid_g$g \
   python tools/bench_sigma_grid.py --child --workload synth-500 --reps 1 > gpurun_out/ncu_grid_g$g.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -c 1 -o gpurun_out/r01_contract_v1 \
   python tools/bench_contract.py --reps 1 --nb 766 --naux 3830 --homo 72 --only epsilon_syrk --out gpurun_out/sweep_ncu_v1.jsonl > gpurun_out/ncu_contract_v1.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c60_v1.json 2> gpurun_out/bench_c60_v1.err
tail -n 3 gpurun_out/bench_c60_v1.err
