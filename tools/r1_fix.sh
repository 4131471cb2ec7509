#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ncu --set full --clock-control none --import-source on -k 'regex:lz4_compress_kernel.*short' -s 3 -c 1 -o gpurun_out/prof_r1_lz4_compress \
      python bench.py --profile --codec lz4 --op compress --steps 1 --warmup 3 --blocks 16384 > gpurun_out/ncu_lz4_compress.log 2>&1
tail -1 gpurun_out/ncu_lz4_compress.log | cut -c1-120
timeout 600 python bench.py --codec zstd --op compress --steps 5 --warmup 3 --blocks 16384 --e2e-steps 2 2>/dev/null | tail -1 > gpurun_out/final_zstd_compress.json; cut -c1-300 gpurun_out/final_zstd_compress.json
python - <<'PY'
import json; d=json.load(open('gpurun_out/final_zstd_compress.json')); print(d['e2e'])
PY
