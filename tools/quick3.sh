#!/bin/bash
for spec in "lz4 compress" "snappy compress"; do
  set -- $spec
  timeout 200 python bench.py --codec $1 --op $2 --steps 3 --warmup 3 --blocks 16384 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('$1 $2', round(d['value'],1), 'GiB/s frac', round(d['roofline']['frac'],4), 'ratio', round(d['config']['ratio'],4))"
done
