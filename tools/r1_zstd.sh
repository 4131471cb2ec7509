#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_zstd.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --codec zstd --op decompress --steps 5 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 2>/dev/null | tail -1 > /tmp/l.json
python -c "import json; d=json.load(open('/tmp/l.json')); print('zstd decompress', round(d['value'],1), 'GiB/s', round(d['ms_per_step'],2))"
prof() { # name codec op kernel-regex blocks
  ncu --set full --clock-control none --import-source on -k regex:$4 -s 3 -c 1 -o gpurun_out/prof_r1b_$1 \
      python bench.py --profile --codec $2 --op $3 --steps 1 --warmup 3 --blocks $5 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-120
}
prof lz4_decompress lz4 decompress lz4_decompress_kernel 65536
prof snappy_decompress snappy decompress snappy_decompress_kernel 32768
prof zstd_decompress zstd decompress zstd_decompress_kernel 8192
