#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_lz_parity.py -x -q -m gpu -k snappy 2>&1 | tail -3
for d in 0 6 7 4; do
  timeout 300 python bench.py --codec snappy --op decompress --steps 5 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 --decoder $d 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('snappy decoder $d', round(d['value'],1), 'GiB/s', round(d['ms_per_step'],2))"
done
for c in 6 8 10 12; do
  timeout 300 python bench.py --codec lz4 --op decompress --steps 5 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 --ctas-per-sm $c 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('lz4 grid ctas/sm $c', round(d['value'],1), 'GiB/s', round(d['ms_per_step'],2))"
done
