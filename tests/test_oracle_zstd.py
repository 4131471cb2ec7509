"""CPU tests: pin the Zstandard oracle (restated Java frame decoder + level-3 compressor) against the
reference's fixtures and known answers, and against the reference's bundled libzstd 1.5.6."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zstd")
R_NOT_ENOUGH, R_CORRUPTED, R_BAD_MAGIC = 32, 35, 36


def _read(name):
    return open(os.path.join(G, name), "rb").read()


def test_fixture_frames_decode(oracle):
    # T/zstd/AbstractTestZstd.java:41-67 (with checksum + output padding; concatenated frames)
    for name in ("with-checksum", "multiple-frames"):
        plain, z = _read(name), _read(name + ".zst")
        assert oracle.decompress("zstd", z, len(plain)) == plain
        assert oracle.decompress("zstd", z, len(plain) + 2042) == plain


def test_fixture_errors(oracle):
    # T/zstd/AbstractTestZstd.java:69-78 "Input is corrupted", :175-184 "Invalid magic prefix"
    r, off, _ = oracle.decompress_raw("zstd", _read("offset-before-start.zst"), 20000)
    assert (-r) & 0xFF == 1 and (-r) >> 8 == R_CORRUPTED
    r, off, _ = oracle.decompress_raw("zstd", _read("bad-second-frame.zst"), len(_read("multiple-frames")))
    assert (-r) >> 8 == R_BAD_MAGIC and off == 4076
    # :186-193 truncated frame
    r, off, _ = oracle.decompress_raw("zstd", bytes([40, 256 - 75, 47, 256 - 3, 32, 0, 1, 0]), 1024)
    assert (-r) >> 8 == R_NOT_ENOUGH
    # :195-242 bad Huffman data
    bad = bytes([0x28, 0xB5, 0x2F, 0xFD, 0, 0, 0xF4, 0, 0, 0x0A, 0, 0x3C, 0, 128, 0x10, 255, 255, 255, 255, 255, 255]) + bytes(10)
    r, off, _ = oracle.decompress_raw("zstd", bad, 10)
    assert r < 0 and (-r) & 0xFF == 1


def test_max_compressed_length_known_answers(oracle):
    # T/zstd/AbstractTestZstd.java:140-147
    assert [oracle.max_compressed_length("zstd", n) for n in (0, 65536, 131072, 131073)] == [64, 65824, 131584, 131585]


def test_frame_header_known_answers(oracle):
    # T/zstd/TestCompressor.java:52-92: header layout for sizes around the descriptor thresholds
    import ctypes as C
    for n in (0, 1, 255, 256, 65791, 65792, 131072, 300000):
        data = bytes((i * 7 + 3) & 0xFF for i in range(n))
        z = oracle.compress("zstd", data)
        assert z[:4] == bytes([0x28, 0xB5, 0x2F, 0xFD])
        fhd = z[4]
        assert fhd & 0x04                                   # checksum flag always set (ZstdFrameCompressor.java:74)
        assert (fhd >> 6) == (n >= 256) + (n >= 65792)      # content size descriptor
        buf = np.frombuffer(z, dtype=np.uint8)
        size = oracle.lib.orc_zstd_decompressed_size(buf.ctypes.data_as(C.POINTER(C.c_uint8)), len(z), None)
        assert size == n                                    # AbstractTestZstd.testGetDecompressedSize :149-173
        assert oracle.decompress("zstd", z, n) == data


def test_roundtrip_matrix_vs_libzstd(oracle, refnative, sample_blocks, synthetic_cases):
    rng = np.random.default_rng(3)
    extra = [_read("incompressible"), b"\x07" * 168890,                                  # large RLE (3-byte header)
             bytes(rng.integers(0, 256, 200000, dtype=np.uint8)) + b"abcabcabd" * 40]      # incompressible literals then small literals
    tin = tout = tref = 0
    for blk in synthetic_cases + sample_blocks + extra:
        c = oracle.compress("zstd", blk)
        assert len(c) <= max(oracle.max_compressed_length("zstd", len(blk)), 64)
        assert refnative.decompress("zstd", c, len(blk)) == blk          # verify decompressor accepts the Java-format stream
        assert oracle.decompress("zstd", c, len(blk)) == blk             # exact-size output
        for lvl in (3, 1):
            c2 = refnative.compress("zstd", blk, lvl)
            assert oracle.decompress("zstd", c2, len(blk)) == blk        # we decode libzstd's streams
            assert oracle.decompress("zstd", c2, len(blk) + 1021) == blk
        tin += len(blk); tout += len(c); tref += len(refnative.compress("zstd", blk, 3))
    assert abs(tout / tin - tref / tin) < 0.02, (tout / tin, tref / tin)   # same DFAST level-3 class ratio


def test_multiblock_and_concatenated_frames(oracle, refnative, pieces):
    data = np.concatenate(pieces[:6]).tobytes()           # 768 KiB -> 6 blocks with a carried window
    z = oracle.compress("zstd", data)
    assert refnative.decompress("zstd", z, len(data)) == data and oracle.decompress("zstd", z, len(data)) == data
    z2 = refnative.compress("zstd", data, 3)
    assert oracle.decompress("zstd", z2 + z, 2 * len(data)) == data + data   # concatenated frames (:56-67)
    r, off, _ = oracle.decompress_raw("zstd", z2, len(data) - 1)
    assert r < 0                                          # output too small


def test_small_literal_roundtrip(oracle, refnative):
    data = bytes(range(256))
    for n in range(1, 256):
        c = oracle.compress("zstd", data[:n])
        assert oracle.decompress("zstd", c, n) == data[:n] and refnative.decompress("zstd", c, n) == data[:n]
