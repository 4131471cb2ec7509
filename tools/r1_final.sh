#!/bin/bash
# round-1 final validation + evidence run (1 GPU)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default)"; timeout 900 python bench.py 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cut -c1-1500 gpurun_out/final_bench.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/final_ref.err | tail -1 > gpurun_out/final_ref.json; cut -c1-600 gpurun_out/final_ref.json
echo "== matrix"; bash tools/bench_matrix.sh gpurun_out/final_matrix.jsonl 32768 2>&1 | tail -8
echo "== xxh64"; timeout 300 python bench.py --codec xxh64 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_xxh64.json; cut -c1-400 gpurun_out/final_xxh64.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --profile --steps 3 --warmup 3 > gpurun_out/final_launches.log 2>&1; grep -c lz4_decompress gpurun_out/final_launches.csv
echo "== ncu full per kernel"; bash tools/profile_all.sh 2>&1 | tail -12
