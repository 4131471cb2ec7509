#!/bin/bash
for k in 2 3 4 8; do
  timeout 200 python bench.py --ctas-per-sm $k --steps 4 --warmup 3 --blocks 32768 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('ctas_per_sm', $k, round(d['value'],1), 'GiB/s frac', round(d['roofline']['frac'],4))"
done
timeout 200 python bench.py --codec xxh64 --steps 5 --warmup 3 --blocks 32768 2>&1 | tail -1 | cut -c1-500
