"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side logic."""
import os
import re

import numpy as np
import pytest

import aircompressor_b200 as acb
from aircompressor_b200 import api, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = acb.lib()
    header = open(os.path.join(ROOT, "include", "aircompress_cuda.h")).read()
    declared = set(re.findall(r"\b(acc_[a-z0-9_]+)\s*\(", header))
    declared -= {"acc_ctx"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/aircompress_cuda.h but not exported"


def test_bounds_match_reference_formulas():
    L = acb.lib()
    assert L.acc_lz4_compress_bound(65536) == 65809            # Lz4RawCompressor.java:64-67
    assert L.acc_snappy_compress_bound(65536) == 76490         # SnappyRawCompressor.java:47-70
    # T/zstd/AbstractTestZstd.java:140-147
    assert [L.acc_zstd_compress_bound(n) for n in (0, 65536, 131072, 131073)] == [64, 65824, 131584, 131585]


def test_no_gpu_means_loud_failure_not_fallback():
    L = acb.lib()
    if L.acc_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        acb.Lz4CudaDecompressor()


def test_verify_range_message():
    arr = np.zeros(10, dtype=np.uint8)
    with pytest.raises(api.IllegalArgumentException, match=r"Invalid offset or length \(5, 6\) in array of length 10"):
        api._verify_range(arr, 5, 6)
    api._verify_range(arr, 5, 5)


def test_snappy_uncompressed_length_host_helper():
    import ctypes
    L = acb.lib()
    buf = np.frombuffer(bytes([0x80, 0x80, 0x04]), dtype=np.uint8)
    assert L.acc_snappy_uncompressed_length(buf.ctypes.data, 3, None) == 65536
    off = ctypes.c_int64(0)
    bad = np.frombuffer(bytes([255, 255, 255, 255, 8]), dtype=np.uint8)
    r = L.acc_snappy_uncompressed_length(bad.ctypes.data, 5, ctypes.byref(off))
    assert (-r) >> 8 == 9  # "invalid compressed length" (T/snappy/AbstractTestSnappy.java:48-56)


def test_partition_by_bytes():
    sizes = np.array([10, 10, 10, 10, 100, 10, 10, 10], dtype=np.int64)
    parts = sharding.partition_by_bytes(sizes, 2)
    assert parts[0][0] == 0 and parts[-1][1] == 8 and parts[0][1] == parts[1][0]
    for w in (1, 2, 3, 4, 8, 16):
        parts = sharding.partition_by_bytes(sizes, w)
        assert len(parts) == w and parts[0][0] == 0 and parts[-1][1] == 8
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
    even = sharding.partition_by_bytes(np.full(65536, 65536), 8)
    assert [e - b for b, e in even] == [8192] * 8


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/aircompress_cuda.h is the contract a cgo / JNI / FFM host binds: it must compile as strict C99 (and as C++),
    and a plain-C consumer (examples/acc_roundtrip.c) must compile warning-free and link against the library.  Without a GPU
    the program must fail loudly at acc_init (exit code 2), never fall back."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    probe = tmp_path / "probe.c"
    probe.write_text('#include "aircompress_cuda.h"\nint main(void) { return acc_lz4_compress_bound(0) == 16 ? 0 : 1; }\n')
    inc = os.path.join(root, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(probe)], check=True)
    subprocess.run([shutil.which("g++") or "/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(probe)], check=True)
    exe = tmp_path / "acc_roundtrip"
    libdir = os.path.join(root, "aircompressor_b200")
    subprocess.run([gcc, "-O2", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", inc, os.path.join(root, "examples", "acc_roundtrip.c"),
                    "-L", libdir, "-laircompress_cuda", "-o", str(exe)], check=True)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe), os.path.join(root, "README.md")], env=dict(os.environ, LD_LIBRARY_PATH=libdir), capture_output=True, text=True)
        assert r.returncode == 2 and "acc_init failed" in r.stderr
