#!/bin/bash
# 1 GPU: parity suite (dense BSE, exact, CDA, new grid + contraction kernels), kernel micro-benchmarks, C60 bench.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -15 gpurun_out/pytest_gpu4.log
XTPB_GRID_GROUP=1 timeout 300 python tools/bench_sigma_grid.py --child --workload synth-1000 --reps 2 > gpurun_out/sigma_grid_v3.jsonl 2>&1
timeout 600 python tools/bench_contract.py --reps 5 --out gpurun_out/contract_sweep_v2.jsonl > gpurun_out/sweep_v2.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c60_v2.json 2> gpurun_out/bench_c60_v2.err
tail -n 3 gpurun_out/bench_c60_v2.err
timeout 600 python bench.py --workload pentacene-tzvp-shape --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pentacene_v2.json 2> gpurun_out/bench_pentacene_v2.err
