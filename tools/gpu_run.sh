#!/bin/bash
# Scratch driver for one gpurun call: runs the steps named on the command line, everything lands in gpurun_out/.
#   tools/gpu_run.sh TAG step [step ...]      steps: tests | multi | trace | c60 | pentacene | benzene | <raw shell after -->
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
NGPU=${NGPU:-1}
if [ "$NGPU" -gt 1 ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29533"
else
  RUN="python"
fi
for step in "$@"; do
  case $step in
    tests)     timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/pytest_$TAG.log ;;
    trace)     XTPB_TRACE=1 XTPB_BENCH_MIN_WARMUP=1 timeout 600 python bench.py --workload c60-tzvp-shape --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/trace_$TAG.json 2> $OUT/trace_$TAG.err; echo "trace rc=$?"; grep "xtpb trace \[" $OUT/trace_$TAG.err | tail -9 ;;
    trace_nosplit) XTPB_TAIL_SPLIT=0 XTPB_TRACE=1 XTPB_BENCH_MIN_WARMUP=1 timeout 600 python bench.py --workload c60-tzvp-shape --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/trace_nosplit_$TAG.json 2> $OUT/trace_nosplit_$TAG.err; echo "trace_nosplit rc=$?"; grep "xtpb trace \[" $OUT/trace_nosplit_$TAG.err | tail -9 ;;
    trace_noovl) XTPB_PPM_OVERLAP=0 XTPB_TRACE=1 XTPB_BENCH_MIN_WARMUP=1 timeout 600 python bench.py --workload c60-tzvp-shape --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/trace_noovl_$TAG.json 2> $OUT/trace_noovl_$TAG.err; echo "trace_noovl rc=$?"; grep "xtpb trace \[" $OUT/trace_noovl_$TAG.err | tail -9 ;;
    diag5)     timeout 600 python tools/bench_contract.py --only aux_rotation,bse_dense_direct_pairs,square_4096_mc_kc,square_4096_mc_mc,bse_direct_step1 --reps 3 --out $OUT/contract_diag5_$TAG.jsonl > $OUT/diag5_$TAG.log 2>&1; echo "diag5 rc=$?"; cut -c1-200 $OUT/diag5_$TAG.log ;;
    contracttests) timeout 600 python -m pytest tests/test_gpu_contract.py -m gpu -x -q > $OUT/pytest_contract_$TAG.log 2>&1; echo "contracttests rc=$?"; tail -4 $OUT/pytest_contract_$TAG.log ;;
    c60)       timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_c60_$TAG.json 2> $OUT/bench_c60_$TAG.err; echo "c60 rc=$?"; cut -c1-600 $OUT/bench_c60_$TAG.json ;;
    c60fast)   timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c60_$TAG.json 2> $OUT/bench_c60_$TAG.err; echo "c60 rc=$?"; cut -c1-600 $OUT/bench_c60_$TAG.json ;;
    pentacene) timeout 600 python bench.py --workload pentacene-tzvp-shape --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_pentacene_$TAG.json 2> $OUT/bench_pentacene_$TAG.err; echo "pentacene rc=$?"; cut -c1-400 $OUT/bench_pentacene_$TAG.json ;;
    multi)     timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/pytest_multi_$TAG.log 2>&1; echo "multi rc=$?"; tail -5 $OUT/pytest_multi_$TAG.log ;;
    sweep)     timeout 1500 python tools/bench_bse_matvec.py --nb 500,1000,2000 --reps 5 --out $OUT/bse_matvec_$TAG.jsonl > $OUT/sweep_$TAG.log 2>&1; echo "sweep rc=$?"; cut -c1-420 $OUT/bse_matvec_$TAG.jsonl ;;
    sweep4000) timeout 1500 python tools/bench_bse_matvec.py --nb 4000 --reps 2 --strategies factorised --out $OUT/bse_matvec4000_$TAG.jsonl > $OUT/sweep4000_$TAG.log 2>&1; echo "sweep4000 rc=$?"; cut -c1-420 $OUT/bse_matvec4000_$TAG.jsonl ;;
    cda)       timeout 1500 python bench.py --workload pentacene-tzvp-cda --sigma cda --evgw 2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_cda_$TAG.json 2> $OUT/bench_cda_$TAG.err; echo "cda rc=$?"; cut -c1-300 $OUT/bench_cda_$TAG.json; tail -3 $OUT/bench_cda_$TAG.err ;;
    c60n)      timeout 900 $RUN bench.py --gpus $NGPU --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c60_n${NGPU}_$TAG.json 2> $OUT/bench_c60_n${NGPU}_$TAG.err; echo "c60 N=$NGPU rc=$?"; cut -c1-200 $OUT/bench_c60_n${NGPU}_$TAG.json ;;
    cdan)      timeout 1500 $RUN bench.py --gpus $NGPU --workload pentacene-tzvp-cda --sigma cda --evgw 2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_cda_n${NGPU}_$TAG.json 2> $OUT/bench_cda_n${NGPU}_$TAG.err; echo "cda N=$NGPU rc=$?"; cut -c1-200 $OUT/bench_cda_n${NGPU}_$TAG.json; tail -3 $OUT/bench_cda_n${NGPU}_$TAG.err ;;
    cpuval)    timeout 1700 python -c "
import json
from oracle import cpu_reference as cr
runs = [cr.validate_against_full_step('pentacene-tzvp-shape', a) for a in ('reference', 'factorised')]
json.dump({'runs': runs}, open('$OUT/cpu_validation_$TAG.json', 'w'), indent=1)
print(json.dumps([(r['algorithm'], r['cores'], r['full_s'], r['sampled_s']) for r in runs]))
" > $OUT/cpuval_$TAG.log 2>&1; echo "cpuval rc=$?"; tail -2 $OUT/cpuval_$TAG.log ;;
    cublas)    timeout 900 python tools/bench_contract.py --cublas --reps 5 --out $OUT/contract_vs_cublas_$TAG.jsonl > $OUT/cublas_$TAG.log 2>&1; echo "cublas rc=$?"; tail -14 $OUT/cublas_$TAG.log | cut -c1-260 ;;
    ncu_eps)   timeout 900 ncu --set full --clock-control none -k regex:contract -c 1 -f -o /tmp/r02_eps_c60 python tools/bench_contract.py --only epsilon_syrk --reps 1 --out $OUT/ncu_eps_$TAG.jsonl > $OUT/ncu_eps_$TAG.log 2>&1; echo "ncu_eps rc=$?"; ncu -i /tmp/r02_eps_c60.ncu-rep --page raw --csv > $OUT/r02_eps_c60_raw.csv 2>/dev/null; ncu -i /tmp/r02_eps_c60.ncu-rep --page details --csv > $OUT/r02_eps_c60_details.csv 2>/dev/null; ls -la $OUT/r02_eps_c60_*.csv ;;
    ncu_rot)   timeout 900 ncu --set full --clock-control none -k regex:contract -c 1 -f -o /tmp/r02_rot_c60 python tools/bench_contract.py --only aux_rotation --reps 1 --out $OUT/ncu_rot_$TAG.jsonl > $OUT/ncu_rot_$TAG.log 2>&1; echo "ncu_rot rc=$?"; ncu -i /tmp/r02_rot_c60.ncu-rep --page raw --csv > $OUT/r02_rot_c60_raw.csv 2>/dev/null; ncu -i /tmp/r02_rot_c60.ncu-rep --page details --csv > $OUT/r02_rot_c60_details.csv 2>/dev/null; ls -la $OUT/r02_rot_c60_*.csv ;;
    launches)  XTPB_BENCH_MIN_WARMUP=1 timeout 1700 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(contract|sigma_ppm|ppm_|chi0_|unpack_|splitk_|symmetrize_|bse_|col_norms|column_dots|residuals_|correction_|olsen_|copy_2d|extract_|scale_|axpby_|scatter_fill|cda_|exact_|set_identity|add_|unit_vectors|gather_strided|cyclic_cols|window_from)' --csv --log-file $OUT/r02_launches_c60.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/launches_$TAG.json 2> $OUT/launches_$TAG.err; echo "launches rc=$?"; wc -l $OUT/r02_launches_c60.csv ;;
    grid)      timeout 900 python tools/bench_sigma_grid.py --workload synth-1000 --out $OUT/sigma_grid_$TAG.jsonl > $OUT/grid_$TAG.log 2>&1; echo "grid rc=$?"; cut -c1-330 $OUT/grid_$TAG.log ;;
    gridtests) timeout 600 python -m pytest tests/test_gpu_gwbse.py -m gpu -x -q -k "grid or ppm or full_gwbse or g0w0" > $OUT/pytest_grid_$TAG.log 2>&1; echo "gridtests rc=$?"; tail -4 $OUT/pytest_grid_$TAG.log ;;
    ncu_grid)  XTPB_SIGMA_GRID=compressed timeout 900 ncu --set full --clock-control none -k regex:sigma_ppm_grid_compressed -c 1 -f -o /tmp/r02_grid python tools/bench_sigma_grid.py --child --workload synth-1000 --reps 1 > $OUT/ncu_grid_$TAG.log 2>&1; echo "ncu_grid rc=$?"; ncu -i /tmp/r02_grid.ncu-rep --page raw --csv > $OUT/r02_grid_raw.csv 2>/dev/null; ncu -i /tmp/r02_grid.ncu-rep --page details --csv > $OUT/r02_grid_details.csv 2>/dev/null; ls -la $OUT/r02_grid_*.csv ;;
    scaletests) timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q > $OUT/pytest_scale_$TAG.log 2>&1; echo "scaletests rc=$?"; tail -15 $OUT/pytest_scale_$TAG.log | cut -c1-300 ;;
    probe)     tools/probe_rcp64 > $OUT/probe_rcp64_$TAG.json 2> $OUT/probe_rcp64_$TAG.err; echo "probe rc=$?"; cat $OUT/probe_rcp64_$TAG.json | cut -c1-200 ;;
    benzene)   timeout 600 python bench.py --workload benzene-tzvp-shape --steps 3 --warmup 3 > $OUT/bench_benzene_$TAG.json 2> $OUT/bench_benzene_$TAG.err; echo "benzene rc=$?"; cut -c1-300 $OUT/bench_benzene_$TAG.json ;;
    walk)      timeout 600 python tools/bench_sigma_grid.py --workload synth-1000 --walk-only --out $OUT/sigma_grid_walk_$TAG.jsonl > $OUT/walk_$TAG.log 2>&1; echo "walk rc=$?"; cut -c1-260 $OUT/walk_$TAG.log ;;
    diag)      timeout 600 python tools/bench_contract.py --only epsilon_syrk,bse_dense_direct_pairs,square_4096_kc_kc,square_4096_mc_kc,square_4096_mc_mc,fill_T_times_Cm --reps 3 --out $OUT/contract_diag_$TAG.jsonl > $OUT/diag_$TAG.log 2>&1; echo "diag rc=$?"; cut -c1-220 $OUT/diag_$TAG.log ;;
    diag0)     XTPB_TAIL_SPLIT=0 timeout 600 python tools/bench_contract.py --only epsilon_syrk --reps 3 --out $OUT/contract_diag0_$TAG.jsonl > $OUT/diag0_$TAG.log 2>&1; echo "diag0 rc=$?"; cut -c1-220 $OUT/diag0_$TAG.log ;;
    ncu_pairs) timeout 900 ncu --set full --clock-control none -k regex:contract -c 1 -f -o /tmp/r02_pairs python tools/bench_contract.py --only bse_dense_direct_pairs --reps 1 --out $OUT/ncu_pairs_$TAG.jsonl > $OUT/ncu_pairs_$TAG.log 2>&1; echo "ncu_pairs rc=$?"; ncu -i /tmp/r02_pairs.ncu-rep --page raw --csv > $OUT/r02_pairs_raw.csv 2>/dev/null; ncu -i /tmp/r02_pairs.ncu-rep --page details --csv > $OUT/r02_pairs_details.csv 2>/dev/null; ls -la $OUT/r02_pairs_*.csv ;;
    smoke)     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log ;;
    *)         echo "unknown step $step" ;;
  esac
done
