#!/bin/bash
# GPU box with 2 GPUs: full parity suite (incl. the 2-rank test), single- and two-rank bench lines.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi2.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -5 gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --workload pentacene-tzvp-shape --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pentacene_n1.json 2> gpurun_out/bench_pentacene_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --workload pentacene-tzvp-shape --steps 2 --warmup 3 > gpurun_out/bench_pentacene_n2.json 2> gpurun_out/bench_pentacene_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_c60_n2.json 2> gpurun_out/bench_c60_n2.err
tail -3 gpurun_out/*.err
