#!/bin/bash
# one `ncu --set full` capture per kernel on the bench workload (1 GPU); reports land in gpurun_out/, summaries via tools/summarize_ncu.py
tag=${1:-r2}
prof() { # name codec op kernel-regex blocks skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -s ${6:-3} -c 1 -o gpurun_out/prof_${tag}_$1 \
      python bench.py --profile --codec $2 --op $3 --steps 1 --warmup 3 --blocks $5 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-120
}
prof lz4_decompress lz4 decompress lz4_decompress_kernel 65536
# launches alternate 16-bit-table / 32-bit-table instantiation (the second exits at once for 64 KiB blocks): skip 4 -> a 16-bit one
prof lz4_compress lz4 compress lz4_compress_kernel 16384 4
prof snappy_decompress snappy decompress snappy_decompress_kernel 32768
prof snappy_compress snappy compress snappy_compress_kernel 16384
prof zstd_decompress zstd decompress zstd_decompress_kernel 8192
prof zstd_compress zstd compress zstd_compress_kernel 8192
prof xxh64 xxh64 decompress xxh64 65536
# every launch of one default bench step with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_lz4_decompress_launches.csv \
    python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
