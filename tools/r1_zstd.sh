#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_zstd.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --codec zstd --op decompress --steps 5 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 2>/dev/null | tail -1 > /tmp/l.json
python -c "import json; d=json.load(open('/tmp/l.json')); print('zstd decompress', round(d['value'],1), 'GiB/s', round(d['ms_per_step'],2))"
