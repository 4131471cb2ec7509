#!/bin/bash
# compute-sanitizer memcheck over the decoder parity tests (corrupted inputs included)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1000 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_lz_parity.py -x -q -m gpu -k "corrupt or overflow or pipelined or matches_oracle" -p no:cacheprovider > gpurun_out/sanitize_lz.log 2>&1; echo "lz rc=$?"; tail -5 gpurun_out/sanitize_lz.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_zstd.py -x -q -m gpu -k "corrupt or fixtures" -p no:cacheprovider > gpurun_out/sanitize_zstd.log 2>&1; echo "zstd rc=$?"; tail -5 gpurun_out/sanitize_zstd.log
