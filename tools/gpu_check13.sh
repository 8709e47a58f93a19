#!/bin/bash
# 1 GPU, last check of the round: whole parity suite incl. the full-size benzene-shape golden test and smoke().
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu13.log
tail -14 gpurun_out/pytest_gpu13.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke13.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke13.log; tail -3 gpurun_out/smoke13.log
