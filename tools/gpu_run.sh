#!/bin/bash
# Scratch driver for one gpurun call: runs the steps named on the command line, everything lands in gpurun_out/.
#   tools/gpu_run.sh TAG step [step ...]      steps: tests | multi | trace | c60 | pentacene | benzene | <raw shell after -->
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for step in "$@"; do
  case $step in
    tests)     timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/pytest_$TAG.log ;;
    trace)     XTPB_TRACE=1 timeout 600 python bench.py --workload c60-tzvp-shape --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/trace_$TAG.json 2> $OUT/trace_$TAG.err; echo "trace rc=$?"; grep "xtpb trace" $OUT/trace_$TAG.err | tail -4 ;;
    c60)       timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_c60_$TAG.json 2> $OUT/bench_c60_$TAG.err; echo "c60 rc=$?"; cut -c1-600 $OUT/bench_c60_$TAG.json ;;
    c60fast)   timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c60_$TAG.json 2> $OUT/bench_c60_$TAG.err; echo "c60 rc=$?"; cut -c1-600 $OUT/bench_c60_$TAG.json ;;
    pentacene) timeout 600 python bench.py --workload pentacene-tzvp-shape --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_pentacene_$TAG.json 2> $OUT/bench_pentacene_$TAG.err; echo "pentacene rc=$?"; cut -c1-400 $OUT/bench_pentacene_$TAG.json ;;
    smoke)     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log ;;
    *)         echo "unknown step $step" ;;
  esac
done
