#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_lz_parity.py tests/test_gpu_block_sizes.py -x -q -m gpu 2>&1 | tail -3
for d in 0 7 5; do
  timeout 300 python bench.py --codec lz4 --op decompress --steps 10 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 0 --decoder $d 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('lz4 decoder $d', round(d['value'],1), 'GiB/s', round(d['ms_per_step'],2))"
done
