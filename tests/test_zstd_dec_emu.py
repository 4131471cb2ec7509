"""CPU test of the Zstandard decode kernel's device code (aircompressor_b200/csrc/zstd_dec.cu).

tests/host/zstd_dec_emu.cpp compiles the SAME source for the host (32 OS threads as the lanes of a warp, barriers as
__syncwarp, an exchange array as shuffles and ballots) and this test compares bytes, lengths, status words and error offsets
with the oracle (= Java decoder rules): frames of the reference algorithm and of libzstd at several levels (multi-block frames
with treeless literals and repeat-mode tables, RLE / raw / predefined modes), long literal runs and long, periodic and far
matches (the ring's slow paths), concatenated frames, corrupted frames, the reference's fixtures, and every output misalignment.
The GPU parity tests remain the gate for the kernel itself.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "zstd")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("zde") / "zstd_dec_emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-DLZS_EMU", "-I" + os.path.join(ROOT, "tests", "host"),
                           "-o", exe, os.path.join(ROOT, "tests", "host", "zstd_dec_emu.cpp")])
    return exe


def run_emu(exe, tmp_path, streams, caps, modes):
    """Runs the emulation once per mode, all modes at the same time (the lanes mostly sleep in barriers); returns {mode: results}."""
    fin = str(tmp_path / "in.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<ii", 0, len(streams)))
        for i, (s, cap) in enumerate(zip(streams, caps)):
            f.write(struct.pack("<qqii", len(s), cap, (i * 7) % 16, (i * 5 + 3) % 16))
            f.write(s)
    procs = {m: subprocess.Popen([exe, fin, str(tmp_path / ("out_%s.bin" % m))] + (["svc"] if m == "svc" else []),
                                 stderr=subprocess.PIPE, text=True) for m in modes}
    res = {}
    for m, p in procs.items():
        _, err = p.communicate(timeout=2400)
        assert p.returncode == 0, (m, p.returncode, err[-400:])
        print(m, err)
        wide, exact = [int(x) for x in __import__("re").findall(r"(\d+) sequences on the wide path, (\d+) in the exact loop", err)[0]]
        assert wide > 4 * exact > 0          # both paths ran, the wide one on most sequences
        out = []
        with open(str(tmp_path / ("out_%s.bin" % m)), "rb") as f:
            for cap in caps:
                olen, status = struct.unpack("<qi", f.read(12))
                out.append((olen, status, f.read(cap + 64)))
        res[m] = out
    return res


def build_cases(oracle, refnative, pieces, synthetic_cases):
    import benchdata
    rng = np.random.default_rng(23)
    blocks = [b.tobytes() for b in benchdata.cut_blocks(pieces, 32 * 1024)[7::260]]
    blocks += [b.tobytes() for b in benchdata.cut_blocks(pieces, 4 * 1024)[::1201]]
    blocks += [s for s in synthetic_cases if len(s) <= 70000]
    noise = bytes(rng.integers(0, 256, 9000, dtype=np.uint8))
    text = pieces[0][:6000].tobytes()
    blocks += [b"head" * 50 + noise + b"x" * 6000 + noise[:5000] + b"tail" * 100,                     # long literals, a long RLE-like match
               text + b"ab" * 3000 + text[:3000] + b"abcdefg" * 500 + text[1000:5000] + noise[:700] + text,   # periodic + far matches between compressible text
               b"\x07" * 40000,                                                                      # RLE block
               np.concatenate(pieces[:2])[:134000].tobytes()]                                       # two blocks per frame: tables carried over
    streams, caps, want = [], [], []
    for i, blk in enumerate(blocks):
        for c in (oracle.compress("zstd", blk), refnative.compress("zstd", blk, 3), refnative.compress("zstd", blk, (1, 9, 19, -5)[i % 4])):
            streams.append(c)
            caps.append(len(blk) + (1021 if i % 3 == 0 else 0))
            want.append(blk)
    streams.append(streams[4] + streams[5]); caps.append(len(want[4]) + len(want[5])); want.append(want[4] + want[5])   # concatenated frames
    n_valid = len(streams)
    for blk in [b for b in blocks if len(b) >= 1000][:8]:
        for c in (oracle.compress("zstd", blk), refnative.compress("zstd", blk, 3)):
            for _ in range(4):
                m = bytearray(c)
                kind = rng.integers(0, 4)
                if kind == 0:
                    m = m[:rng.integers(1, len(m))]
                elif kind == 1:
                    for _k in range(rng.integers(1, 3)):
                        m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)
                elif kind == 2:
                    m[rng.integers(0, min(len(m), 48))] = rng.integers(0, 256)
                streams.append(bytes(m))
                caps.append(len(blk) if kind != 3 else int(rng.integers(0, len(blk))))
    rd = lambda name: open(os.path.join(G, name), "rb").read()
    streams += [rd("with-checksum.zst"), rd("multiple-frames.zst"), rd("offset-before-start.zst"), rd("bad-second-frame.zst"),
                bytes([40, 181, 47, 253, 32, 0, 1, 0])]
    caps += [len(rd("with-checksum")) + 2042, len(rd("multiple-frames")), 20000, len(rd("multiple-frames")), 1024]
    return streams, caps, want, n_valid


def test_decode_device_code_matches_oracle(emu, tmp_path, oracle, refnative, pieces, synthetic_cases):
    # svc: the service kernel's roles (worker warps + the chain warp of a CTA talking through mailboxes); warp: one warp per input
    streams, caps, want, n_valid = build_cases(oracle, refnative, pieces, synthetic_cases)
    expect = [oracle.decompress_raw("zstd", s, cap) for s, cap in zip(streams, caps)]
    for mode, results in run_emu(emu, tmp_path, streams, caps, ["svc", "warp"]).items():
        n_bad = n_reason_diff = 0
        for i, (s, cap) in enumerate(zip(streams, caps)):
            olen, status, data = results[i]
            r, off, ref = expect[i]
            if i < n_valid:
                assert r == len(want[i])
            if r >= 0:
                assert status == 0 and olen == r, (mode, i, hex(status), olen, r)
                assert data[:r] == ref[:r].tobytes(), (mode, i)
            else:
                n_bad += 1
                assert status != 0 and (status & 0xFF) == ((-r) & 0xFF), (mode, i, hex(status), hex(-r))
                if status == -r:
                    assert olen == off, (mode, i, hex(status), olen, off)
                else:
                    n_reason_diff += 1
            assert data[cap:] == b"\xa5" * 64, (mode, i)          # nothing past maxOutputLength is touched
        assert n_bad > 10
        assert n_reason_diff <= max(1, n_bad // 50), (mode, n_reason_diff, n_bad)
