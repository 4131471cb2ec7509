"""CPU test of the record path of the LZ4 / Snappy decoders (aircompressor_b200/csrc/lz_records.cuh + both parse sides + the step
decoders it resumes).

tests/host/lzs_emu.cpp compiles the SAME device source for the host (OS threads as lanes, barriers as __syncwarp, an
exchange array as shuffles and ballots), runs the parse stage and the execute stage over a file of blocks, and this test
compares bytes, lengths, status words and error offsets with the oracle (= Java decoder rules) for valid streams of both
compressors, corrupted streams, the reference's malformed vectors and every (input, output) misalignment class -- once with
roomy record rows and once with rows of 64 records, so that the resume-in-the-middle path runs on every block.  The GPU
parity tests remain the gate for the kernels themselves.
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


def run_emu(exe, tmp_path, codec, streams, caps, row):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<ii", 0 if codec == "lz4" else 1, len(streams)))
        for i, (s, cap) in enumerate(zip(streams, caps)):
            f.write(struct.pack("<qqii", len(s), cap, (i * 7) % 16, (i * 5 + 3) % 16))
            f.write(s)
    subprocess.run([exe, fin, fout, str(row)], check=True, timeout=900)
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


@pytest.mark.parametrize("row", [8192, 64])
@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_record_path_matches_oracle(emu, tmp_path, oracle, refnative, codec, pieces, synthetic_cases, row):
    import benchdata
    rng = np.random.default_rng(7)
    # the emulation pays a thread barrier per shuffle, so the step decoder (the resume path) is slow here: rows of 64 records
    # (= almost everything runs in the resume path) get small blocks only
    if row >= 1024:
        blocks = [b.tobytes() for b in benchdata.cut_blocks(pieces, 64 * 1024)[40::400]]
        blocks += [b.tobytes() for b in benchdata.cut_blocks(pieces, 4 * 1024)[::401]]
        blocks += synthetic_cases
        # long runs and long matches in the middle of a block (literal pieces, periodic and far matches, length extensions)
        noise = bytes(rng.integers(0, 256, 9000, dtype=np.uint8))
        blocks += [b"head" * 50 + noise + b"x" * 6000 + noise[:5000] + b"tail" * 100, noise[:3000] + b"ab" * 4000 + noise[:3000] + b"abc" * 3000 + noise[3000:7000]]
    else:
        blocks = [b.tobytes() for b in benchdata.cut_blocks(pieces, 4 * 1024)[::500]]
        blocks += [s for s in synthetic_cases if len(s) < 6000]
        noise = bytes(rng.integers(0, 256, 700, dtype=np.uint8))
        blocks += [b"head" * 50 + noise + b"x" * 600 + noise[:500] + b"tail" * 100]
    streams, caps = [], []
    for i, blk in enumerate(blocks):
        for comp in (oracle, refnative):
            streams.append(comp.compress(codec, blk))
            caps.append(len(blk) + (1021 if i % 3 == 0 else 0))
    # corrupted streams
    for blk in blocks[:6]:
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
    results = run_emu(emu, tmp_path, codec, streams, caps, row)
    n_bad = check(codec, oracle, streams, caps, results)
    assert n_bad > 3
