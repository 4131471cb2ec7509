#!/bin/bash
# round 2: bench + one ncu --set full capture of the streaming decode kernels (LZ4, Snappy); outputs in gpurun_out/
tag=${1:-s}
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_${tag}_lz4d.json 2> gpurun_out/r2_${tag}_lz4d.err
python - <<PY
import json
for c in ("lz4d","snd"):
    try:
        d=[json.loads(l) for l in open("gpurun_out/r2_${tag}_%s.json"%c) if l.startswith("{")][-1]
        print(c, round(d["value"],1), "GiB/s", round(d["roofline"]["kernel_ms_avg"],2), "ms")
    except Exception as e: print(c, "n/a", e)
PY
timeout 200 python bench.py --codec snappy --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_${tag}_snd.json 2> gpurun_out/r2_${tag}_snd.err
python - <<PY
import json
for c in ("snd",):
    try:
        d=[json.loads(l) for l in open("gpurun_out/r2_${tag}_%s.json"%c) if l.startswith("{")][-1]
        print(c, round(d["value"],1), "GiB/s", round(d["roofline"]["kernel_ms_avg"],2), "ms")
    except Exception as e: print(c, "n/a", e)
PY
if [ "$2" != "noprof" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lz4_stream -s 3 -c 1 -o gpurun_out/prof_r2_${tag}_lz4d \
    python bench.py --profile --steps 1 --warmup 3 --blocks 16384 > gpurun_out/ncu_${tag}_lz4d.log 2>&1
tail -2 gpurun_out/ncu_${tag}_lz4d.log | cut -c1-200
fi
