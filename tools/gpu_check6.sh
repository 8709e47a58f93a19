#!/bin/bash
# 2 GPUs: parity suite incl. TMA instance and the 2-rank test, contraction sweep (TMA vs cp.async), N=1 and N=2 bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contract.py -m gpu -x -q > gpurun_out/pytest_gpu6a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6a.log
tail -8 gpurun_out/pytest_gpu6a.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_contract.py > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
tail -8 gpurun_out/pytest_gpu6.log
timeout 600 python tools/bench_contract.py --reps 5 --out gpurun_out/contract_sweep_v3_tma.jsonl > gpurun_out/sweep_v3.log 2>&1
XTPB_TMA=0 timeout 600 python tools/bench_contract.py --reps 5 --out gpurun_out/contract_sweep_v3_cpasync.jsonl > gpurun_out/sweep_v3b.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c60_v4.json 2> gpurun_out/bench_c60_v4.err
tail -n 3 gpurun_out/bench_c60_v4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_c60_v4_n2.json 2> gpurun_out/bench_c60_v4_n2.err
tail -n 3 gpurun_out/bench_c60_v4_n2.err
