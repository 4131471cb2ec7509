#!/bin/bash
# refresh of the headline evidence after the last LZ4 decode change (1 GPU)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (default)"; timeout 900 python bench.py 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cut -c1-400 gpurun_out/final_bench.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --profile --steps 3 --warmup 3 > gpurun_out/final_launches.log 2>&1; grep -c lz4_decompress gpurun_out/final_launches.csv
echo "== ncu full lz4 decompress"
ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_kernel -s 3 -c 1 -o gpurun_out/prof_r1_lz4_decompress python bench.py --profile --codec lz4 --op decompress --steps 1 --warmup 3 --blocks 65536 > gpurun_out/ncu_lz4_decompress.log 2>&1; tail -1 gpurun_out/ncu_lz4_decompress.log | cut -c1-100
