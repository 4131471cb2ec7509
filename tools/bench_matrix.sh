#!/bin/bash
# runs bench.py for every codec/direction (device-resident throughput) and collects the JSON lines
out=${1:-gpurun_out/matrix.jsonl}
blocks=${2:-16384}
: > $out
for spec in "lz4 decompress" "lz4 compress" "snappy decompress" "snappy compress" "zstd decompress" "zstd compress"; do
  set -- $spec
  b=$blocks; if [ "$1" = "zstd" ]; then b=$((blocks/2)); fi
  timeout 600 python bench.py --codec $1 --op $2 --steps 5 --warmup 3 --blocks $b --e2e-steps 2 2>/dev/null | tail -1 >> $out
done
python - <<'PY' $out
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:100]); continue
    r=d["roofline"]; c=d.get("cpu_baseline") or {}; e=d.get("e2e") or {}
    print(f'{d["metric"]:45s} {d["value"]:9.1f} GiB/s  frac={r["frac"]:.4f} frac_min_traffic={r["frac_min_traffic"]:.4f} ratio={d["config"]["ratio"]:.3f} cpu={c.get("value",0):.2f} GiB/s x{c.get("cores")} e2e={e.get("value")}')
PY
