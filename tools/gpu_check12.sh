#!/bin/bash
# 4 GPUs: C60-shape bench line at N=4 (the driver's round-end scaling run covers 1/2/4/8).
set -x
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c60_n4_r12.json 2> gpurun_out/bench_c60_n4_r12.err
tail -n 3 gpurun_out/bench_c60_n4_r12.err; head -c 300 gpurun_out/bench_c60_n4_r12.json
