#!/bin/bash
for k in 64 256 1024; do
  timeout 200 python bench.py --decoder 2 --tpb-ctas $k --steps 3 --warmup 3 --blocks 65536 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('tpb ctas', $k, round(d['value'],1), 'GiB/s frac', round(d['roofline']['frac'],4))"
done
