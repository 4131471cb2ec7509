#!/bin/bash
# block-size sweep (BASELINE.json configs[4]) on one GPU: device-resident throughput of the three decoders and encoders
cd "${GRAFT_REPO_ROOT:-/root/repo}"
out=gpurun_out/block_sweep.jsonl; : > $out
for kib in 4 16 64 256 1024; do
  total=$((4*1024*1024)); [ $kib -le 16 ] && total=$((2*1024*1024))   # KiB of uncompressed data per run
  blocks=$((total/kib))
  for spec in "lz4 decompress" "snappy decompress" "zstd decompress" "lz4 compress" "snappy compress" "zstd compress"; do
    set -- $spec
    b=$blocks; [ "$2" = "compress" ] && b=$((blocks/4))
    timeout 300 python bench.py --codec $1 --op $2 --block-kib $kib --blocks $b --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 0 2>/dev/null | tail -1 >> $out
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/block_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad', l[:80]); continue
    print(d['config']['workload'][:60], round(d['value'],1), 'GiB/s ratio', round(d['config']['ratio'],3))
PY
