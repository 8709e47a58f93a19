#!/bin/bash
# GPU box: parity tests, a short bench, the ncu launch list and full captures of the two top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload pentacene-tzvp-shape --steps 2 --warmup 3 > gpurun_out/bench_pentacene.json 2> gpurun_out/bench_pentacene.err
XTPB_BENCH_MIN_WARMUP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
   --log-file gpurun_out/launches_pentacene.csv python bench.py --workload pentacene-tzvp-shape --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -c 4 -o gpurun_out/r01_contract \
   python tools/bench_contract.py --reps 1 --nb 766 --naux 3830 --homo 72 --only epsilon_syrk,aux_rotation --out gpurun_out/sweep_ncu.jsonl > gpurun_out/ncu_contract.log 2>&1
timeout 900 python tools/bench_contract.py --reps 5 --out gpurun_out/contract_sweep.jsonl > gpurun_out/sweep.log 2>&1
ls -la gpurun_out
