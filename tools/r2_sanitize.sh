#!/bin/bash
# compute-sanitizer over decoder AND encoder tests: memcheck, racecheck (shared-memory hazards: the encoders' tables, ballots and
# atomics on staged bitstreams), initcheck (reads of uninitialised device memory).  Summaries land in gpurun_out/.
SEL="tests/test_gpu_lz_parity.py::test_decompress_error_parity_on_corrupt_streams tests/test_gpu_lz_parity.py::test_compress_roundtrips_through_reference_decoders tests/test_gpu_zstd.py::test_decode_error_parity_on_corrupt_frames tests/test_gpu_zstd.py::test_compress_roundtrips_through_reference_decoders tests/test_gpu_xxh64.py"
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_$tool.log python -m pytest $SEL -x -q > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "== $tool: pytest: $(tail -1 gpurun_out/sanitizer_${tool}_pytest.log)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized" gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -8
done
