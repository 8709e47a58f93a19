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
    multi)     timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/pytest_multi_$TAG.log 2>&1; echo "multi rc=$?"; tail -5 $OUT/pytest_multi_$TAG.log ;;
    sweep)     timeout 1500 python tools/bench_bse_matvec.py --nb 500,1000,2000 --reps 5 --out $OUT/bse_matvec_$TAG.jsonl > $OUT/sweep_$TAG.log 2>&1; echo "sweep rc=$?"; tail -3 $OUT/sweep_$TAG.log ;;
    sweep4000) timeout 1500 python tools/bench_bse_matvec.py --nb 4000 --reps 2 --strategies factorised --out $OUT/bse_matvec4000_$TAG.jsonl > $OUT/sweep4000_$TAG.log 2>&1; echo "sweep4000 rc=$?"; tail -4 $OUT/sweep4000_$TAG.log ;;
    cda)       timeout 1500 python bench.py --workload pentacene-tzvp-cda --sigma cda --evgw 2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_cda_$TAG.json 2> $OUT/bench_cda_$TAG.err; echo "cda rc=$?"; cut -c1-300 $OUT/bench_cda_$TAG.json; tail -3 $OUT/bench_cda_$TAG.err ;;
    smoke)     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log ;;
    *)         echo "unknown step $step" ;;
  esac
done
