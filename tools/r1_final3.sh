#!/bin/bash
# last confirmation of the committed tree (1 GPU): tests, smoke, headline bench, then the codec matrix
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (default)"; timeout 600 python bench.py 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench.json
echo "== matrix"; : > gpurun_out/final_matrix2.jsonl
for spec in "zstd compress 16384" "lz4 decompress 32768" "snappy decompress 32768" "zstd decompress 16384" "lz4 compress 32768" "snappy compress 32768"; do
  set -- $spec
  timeout 300 python bench.py --codec $1 --op $2 --steps 5 --warmup 3 --blocks $3 --e2e-steps 2 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 >> gpurun_out/final_matrix2.jsonl
  tail -1 gpurun_out/final_matrix2.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value'],1), 'ratio', round(d['config']['ratio'],4), 'e2e', d['e2e'].get('value'))"
done
