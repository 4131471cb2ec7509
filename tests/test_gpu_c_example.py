"""A plain-C host (examples/acc_roundtrip.c) drives the C ABI end to end on the GPU: one compress batch, one decompress batch,
pinned staging from acc_host_alloc -- the call sequence a cgo / JNI / FFM binding makes, with no Python in between."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
def test_c_example_round_trips(tmp_path, codec):
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    exe = tmp_path / "acc_roundtrip"
    libdir = os.path.join(ROOT, "aircompressor_b200")
    subprocess.run([gcc, "-O2", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "acc_roundtrip.c"),
                    "-L", libdir, "-laircompress_cuda", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), os.path.join(ROOT, "tests", "golden", "silesia_sample.bin"), codec],
                       env=dict(os.environ, LD_LIBRARY_PATH=libdir), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "round trip ok" in r.stdout
