#!/bin/bash
# round-1 check of the pipelined host path + xxh64 ncu capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_lz_parity.py -x -q -m gpu -k "pipelined or full_batch or java_shaped" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_xxh64.py -x -q -m gpu 2>&1 | tail -3
for spec in "lz4 decompress" "lz4 compress" "snappy decompress" "zstd decompress"; do
  set -- $spec
  timeout 400 python bench.py --codec $1 --op $2 --steps 5 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 3 2>gpurun_out/pipe_$1_$2.err | tail -1 > gpurun_out/pipe_$1_$2.json
  python -c "import json; d=json.load(open('gpurun_out/pipe_$1_$2.json')); print('$1 $2', round(d['value'],1), 'GiB/s e2e', d['e2e'])"
done
for k in 2 4 16; do
  timeout 300 python bench.py --codec lz4 --op decompress --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 3 --pipeline $k 2>/dev/null | tail -1 > /tmp/l.json
  python -c "import json; d=json.load(open('/tmp/l.json')); print('pipeline $k e2e', d['e2e']['value'])"
done
ncu --set full --clock-control none --import-source on -k regex:xxh64 -s 3 -c 1 -o gpurun_out/prof_r1_xxh64 \
      python bench.py --profile --codec xxh64 --steps 1 --warmup 3 --blocks 65536 > gpurun_out/ncu_xxh64.log 2>&1
tail -1 gpurun_out/ncu_xxh64.log | cut -c1-160
