#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== block size sweep tests"; timeout 900 python -m pytest tests/test_gpu_block_sizes.py -x -q -m gpu 2>&1 | tail -4
echo "== lz4 compress capture"
ncu --set full --clock-control none --import-source on -k regex:lz4_compress_kernel -s 4 -c 1 -o gpurun_out/prof_r1_lz4_compress python bench.py --profile --codec lz4 --op compress --steps 1 --warmup 3 --blocks 16384 > gpurun_out/ncu_lz4_compress.log 2>&1; tail -1 gpurun_out/ncu_lz4_compress.log | cut -c1-100
ls corpus/silesia 2>/dev/null | head -3
echo "== full-corpus matrix"
for spec in "lz4 decompress 65536" "snappy decompress 65536" "zstd decompress 32768" "lz4 compress 32768" "snappy compress 32768" "zstd compress 16384"; do
  set -- $spec
  timeout 600 python bench.py --codec $1 --op $2 --steps 5 --warmup 3 --blocks $3 --e2e-steps 2 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 >> gpurun_out/fullcorpus_matrix.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/fullcorpus_matrix.jsonl'):
    d=json.loads(l); print(d['metric'], round(d['value'],1), 'ratio', round(d['config']['ratio'],4), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e'].get('value'), d['data'][:40])
PY
