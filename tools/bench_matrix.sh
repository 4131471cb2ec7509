#!/bin/bash
# bench.py for every codec/direction at the full batch (4 GiB of uncompressed data per GPU), with cpu_baseline and e2e; collects the JSON lines
out=${1:-gpurun_out/matrix.jsonl}
: > $out
for spec in "lz4 decompress" "lz4 compress" "snappy decompress" "snappy compress" "zstd decompress" "zstd compress" "xxh64 hash"; do
  set -- $spec
  timeout 900 python bench.py --codec $1 --op $2 --steps 5 --warmup 3 --e2e-steps 5 --no-extra 2>/dev/null | tail -1 >> $out
done
python - <<'PY' $out
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:100]); continue
    r=d["roofline"]; c=d.get("cpu_baseline") or {}; e=d.get("e2e") or {}; nat=(c.get("reference_native") or {})
    print(f'{d["metric"]:42s} {d["value"]:8.1f} GiB/s frac={r["frac"]:.4f} (U+C {r["frac_min_traffic"]:.4f}) ratio={d["config"]["ratio"] or 0:.3f} cpu port={c.get("value") or 0:.2f} native={nat.get("value") or 0:.2f} x{c.get("cores")} e2e={e.get("value")}')
PY
