#!/bin/bash
# 1 GPU: whole parity suite, timings of the two grid-scan variants, C60 bench line, ncu --set full of the compressed-scan
# kernels, complete launch list of a pentacene-shape step (a C60-shape step is too slow under ncu, see gpu_check8).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu9.log
tail -12 gpurun_out/pytest_gpu9.log
timeout 200 python tools/bench_sigma_grid.py --workload synth-1000 --reps 2 --out gpurun_out/sigma_grid9.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c60_r9.json 2> gpurun_out/bench_c60_r9.err
tail -n 3 gpurun_out/bench_c60_r9.err; head -c 400 gpurun_out/bench_c60_r9.json
XTPB_SIGMA_GRID=compressed timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:grid_compressed|ppm_moments' -c 2 -o gpurun_out/r01_sigma_grid_compressed_v2 \
   python tools/bench_sigma_grid.py --child --workload synth-500 --reps 1 > gpurun_out/ncu_grid_compressed_v2.log 2>&1
XTPB_BENCH_MIN_WARMUP=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
   --log-file gpurun_out/launches_pentacene_r9.csv python bench.py --workload pentacene-tzvp-shape --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_pentacene_r9.log 2>&1
timeout 200 python bench.py --workload pentacene-tzvp-shape --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pentacene_r9.json 2> gpurun_out/bench_pentacene_r9.err
ls -la gpurun_out | tail -12
