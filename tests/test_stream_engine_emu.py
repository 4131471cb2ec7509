"""CPU test of the two-phase LZ77 decode engine (aircompressor_b200/csrc/lz_stream.cuh + the LZ4 / Snappy parse sides).

tests/host/lzs_emu.cpp compiles the SAME device source for the host (OS threads as lanes, barriers as __syncwarp, an
exchange array as shuffles and ballots) and decodes a file of blocks; this test compares bytes, lengths, status words and
error offsets with the oracle (= Java decoder rules) for valid streams of both compressors, corrupted streams, the
reference's malformed vectors and every (input, output) misalignment class.  The GPU parity tests remain the gate for
the kernels themselves; this one checks the parse logic and the parse -> execute hand-over on the CPU.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lzs") / "lzs_emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-DLZS_EMU", "-I" + os.path.join(ROOT, "tests", "host"),
                           "-o", exe, os.path.join(ROOT, "tests", "host", "lzs_emu.cpp")])
    return exe


def run_emu(exe, tmp_path, codec, streams, caps):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<ii", 0 if codec == "lz4" else 1, len(streams)))
        for i, (s, cap) in enumerate(zip(streams, caps)):
            f.write(struct.pack("<qqii", len(s), cap, (i * 7) % 16, (i * 5 + 3) % 16))
            f.write(s)
    subprocess.run([exe, fin, fout], check=True, timeout=600)
    out = []
    with open(fout, "rb") as f:
        for cap in caps:
            olen, status = struct.unpack("<qi", f.read(12))
            out.append((olen, status, f.read(cap + 64)))
    return out


def check(codec, oracle, streams, caps, results):
    n_bad = 0
    for i, (s, cap) in enumerate(zip(streams, caps)):
        olen, status, data = results[i]
        r, off, ref_out = oracle.decompress_raw(codec, s, cap)
        if r >= 0:
            assert status == 0 and olen == r, (i, hex(status), olen, r)
            assert data[:r] == ref_out[:r].tobytes(), i
        else:
            n_bad += 1
            assert status == -r and olen == off, (i, hex(status), hex(-r), olen, off)
        assert data[cap:] == b"\xa5" * 64, i          # nothing past maxOutputLength is touched
    return n_bad


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_stream_engine_matches_oracle(emu, tmp_path, oracle, refnative, codec, pieces, synthetic_cases):
    import benchdata
    rng = np.random.default_rng(7)
    blocks = [b.tobytes() for b in benchdata.cut_blocks(pieces, 64 * 1024)[::29]]
    blocks += [b.tobytes() for b in benchdata.cut_blocks(pieces, 4 * 1024)[::401]]
    blocks += [b.tobytes() for b in benchdata.cut_blocks(pieces, 128 * 1024)[::61]]
    blocks += synthetic_cases
    # long runs and long matches in the middle of a block (literal pieces, periodic and far matches, length extensions)
    noise = bytes(rng.integers(0, 256, 9000, dtype=np.uint8))
    blocks += [b"head" * 50 + noise + b"x" * 6000 + noise[:5000] + b"tail" * 100, noise[:3000] + b"ab" * 4000 + noise[:3000] + b"abc" * 3000 + noise[3000:7000]]
    streams, caps = [], []
    for i, blk in enumerate(blocks):
        for comp in (oracle, refnative):
            streams.append(comp.compress(codec, blk))
            caps.append(len(blk) + (1021 if i % 3 == 0 else 0))
    # corrupted streams
    for blk in blocks[:10]:
        c = bytearray(oracle.compress(codec, blk))
        if len(c) < 8:
            continue
        for _ in range(5):
            m = bytearray(c)
            kind = rng.integers(0, 4)
            if kind == 0:
                m = m[:rng.integers(1, len(m))]
            elif kind == 1:
                for _k in range(rng.integers(1, 4)):
                    m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)
            elif kind == 2:
                m[rng.integers(0, min(len(m), 64))] = rng.integers(0, 256)
            streams.append(bytes(m))
            caps.append(len(blk) if kind != 3 else int(rng.integers(0, len(blk))))
    if codec == "lz4":
        streams += [bytes([15, 0, 0, 255, 255, 138, 49, 255, 255, 0]), b"", b"\x00", b"\x00", b"\x10"]
        caps += [1024, 16, 0, 5, 0]
    else:
        streams += [bytes([16, 1, 0, 1, 0, 1, 0, 1, 0]), bytes([128, 8, 252, 255, 255, 255, 127, 0, 0, 0, 0, 0, 0, 0, 0]),
                    bytes([255, 255, 255, 255, 8]), b"", bytes([0x80])]
        caps += [64, 1024, 64, 8, 8]
    results = run_emu(emu, tmp_path, codec, streams, caps)
    n_bad = check(codec, oracle, streams, caps, results)
    assert n_bad > 5
