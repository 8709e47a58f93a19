#!/bin/bash
# 2 GPUs: parity suite (deferred metric rotation, block Gram-Schmidt, 2-rank test), N=1 and N=2 bench lines.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
tail -15 gpurun_out/pytest_gpu5.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c60_v3.json 2> gpurun_out/bench_c60_v3.err
tail -n 3 gpurun_out/bench_c60_v3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_c60_v3_n2.json 2> gpurun_out/bench_c60_v3_n2.err
tail -n 3 gpurun_out/bench_c60_v3_n2.err
