#!/bin/bash
# 2 GPUs: sharded pipeline against the oracle (tests/test_gpu_multi.py), then a C60-shape bench line at N=2.
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi11.log
tail -8 gpurun_out/pytest_multi11.log
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c60_n2_r11.json 2> gpurun_out/bench_c60_n2_r11.err
tail -n 3 gpurun_out/bench_c60_n2_r11.err; head -c 300 gpurun_out/bench_c60_n2_r11.json
