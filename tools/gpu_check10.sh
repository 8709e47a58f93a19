#!/bin/bash
# 1 GPU: whole parity suite (incl. the GWBSE driver test), grid-scan variants (register caps), C60 bench line with the
# fastest variant.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10.log
tail -12 gpurun_out/pytest_gpu10.log
timeout 300 python tools/bench_sigma_grid.py --workload synth-1000 --reps 2 --out gpurun_out/sigma_grid10.jsonl
best=$(python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/sigma_grid10.jsonl")]
ref = [r for r in rows if r.get("mode") == "direct" and "checksum" in r]
ok = [r for r in rows if r.get("mode") == "compressed" and "checksum" in r and ref and abs(r["checksum"] - ref[0]["checksum"]) < 1e-9 * abs(ref[0]["checksum"])]
print(min(ok, key=lambda r: r["ms"])["min_blocks_per_sm"] if ok else "5")
PY
)
echo "best occupancy cap: $best" | tee gpurun_out/grid_best10.txt
XTPB_GRID_OCC=$best timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c60_r10.json 2> gpurun_out/bench_c60_r10.err
tail -n 3 gpurun_out/bench_c60_r10.err; head -c 300 gpurun_out/bench_c60_r10.json
ls -la gpurun_out | tail -6
