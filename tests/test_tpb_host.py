"""CPU test: the thread-per-block LZ4 decoder (aircompressor_b200/csrc/lz4_tpb.cuh) is plain scalar code that also
compiles for the host.  tests/host/lz4_tpb_host_test.cpp runs it against the oracle on ~170,000 cases (valid, truncated,
bit-flipped, short-capacity streams; every input/output alignment) with guard bands around the output."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tpb_decoder_matches_oracle_on_host(tmp_path):
    exe = tmp_path / "lz4_tpb_host_test"
    cmd = ["g++", "-O2", "-I", os.path.join(ROOT, "aircompressor_b200", "csrc"), "-o", str(exe),
           os.path.join(ROOT, "tests", "host", "lz4_tpb_host_test.cpp"), os.path.join(ROOT, "oracle", "liboracle.so"),
           "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)
    out = subprocess.run([str(exe), os.path.join(ROOT, "tests", "golden", "silesia_sample.bin")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-500:]
    # "tests N ok A fallback B (valid streams falling back C, invalid D)"
    words = out.stdout.split()
    n_tests, n_ok = int(words[1]), int(words[3])
    valid_fallback = int(words[10].rstrip(","))
    assert n_tests > 100000 and n_ok > 50000
    assert valid_fallback < 100   # the optimistic decoder handles essentially every valid stream itself
