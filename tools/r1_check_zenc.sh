#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_zstd.py tests/test_gpu_block_sizes.py -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --codec zstd --op compress --steps 5 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 2 2>/dev/null | tail -1 > /tmp/l.json
python -c "import json; d=json.load(open('/tmp/l.json')); print('zstd compress', round(d['value'],1), 'GiB/s ratio', round(d['config']['ratio'],4), 'e2e', d['e2e'].get('value'), d['e2e'].get('error'))"
