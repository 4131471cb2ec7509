"""CPU tests: pin the oracle (restated Java algorithms) against the reference's own golden vectors
and against the reference's bundled native libraries (oracle/_ref)."""
import numpy as np
import pytest

PRIME = 2654435761
PRIME32 = 0x9E3779B1
PRIME64 = 0x9E3779B185EBCA8D
M64 = (1 << 64) - 1


def _zstd_test_buffer():
    # T/zstd/TestXxHash64.java:30-38
    buf = bytearray(101)
    v = PRIME
    for i in range(101):
        buf[i] = (v >> 24) & 0xFF
        v = (v * v) & M64
    return bytes(buf)


def test_xxh64_known_answers_zstd(oracle):
    # T/zstd/TestXxHash64.java:41-60
    buf = _zstd_test_buffer()
    exp = [
        (0, 0, 0xEF46DB3751D8E999), (0, 1, 0x4FCE394CC88952D8), (PRIME, 1, 0x739840CB819FA723),
        (0, 4, 0x9256E58AA397AEF1), (PRIME, 4, 0x9D5FFDFB928AB4B), (0, 8, 0xF74CB1451B32B8CF),
        (PRIME, 8, 0x9C44B77FBCC302C5), (0, 14, 0xCFFA8DB881BC3A3D), (PRIME, 14, 0x5B9611585EFCC9CB),
        (0, 32, 0xAF5753D39159EDEE), (PRIME, 32, 0xDCAB9233B8CA7B0F), (0, 101, 0x0EAB543384F878AD),
        (PRIME, 101, 0xCAA65939306F1E21),
    ]
    for seed, n, want in exp:
        assert oracle.xxh64(buf[:n], seed) == want, (seed, n)


def _sanity_buffer(n):
    # T/xxhash/AbstractTestXxHash64.java:51-60
    out = bytearray(n)
    g = PRIME32
    for i in range(n):
        out[i] = (g >> 56) & 0xFF
        g = (g * PRIME64) & M64
    return bytes(out)


def test_xxh64_official_sanity_vectors(oracle):
    # T/xxhash/AbstractTestXxHash64.java:65-92
    assert oracle.xxh64(b"", 0) == 0xEF46DB3751D8E999
    assert oracle.xxh64(b"", PRIME32) == 0xAC75FDA2929B17EF
    for n, seed, want in [(1, 0, 0xE934A84ADB052768), (1, PRIME32, 0x5014607643A9B4C3), (4, 0, 0x9136A0DCA57457EE),
                          (14, 0, 0x8282DCC4994E35C8), (14, PRIME32, 0xC3BD6BF63DEB6DF0), (222, 0, 0xB641AE8CB691C174),
                          (222, PRIME32, 0x20CB8AB7AE10C14A)]:
        assert oracle.xxh64(_sanity_buffer(n), seed) == want, (n, seed)


def test_xxh64_vs_native_many_lengths(oracle, refnative):
    # T/zstd/TestXxHash64.java:63-72 (all-zero inputs of every length) + random data
    rng = np.random.default_rng(7)
    data = bytes(rng.integers(0, 256, 4096, dtype=np.uint8))
    for n in list(range(0, 300)) + [1023, 1024, 1025, 4096]:
        assert oracle.xxh64(bytes(n), 0) == refnative.xxh64(bytes(n), 0)
        assert oracle.xxh64(data[:n], 12345) == refnative.xxh64(data[:n], 12345)


def test_lz4_known_answer_errors(oracle):
    # T/lz4/TestLz4.java:53-60 -> "offset outside destination buffer: offset=3"
    r, off, _ = oracle.decompress_raw("lz4", bytes([15, 0, 0, 255, 255, 138, 49, 255, 255, 0]), 1024)
    assert (-r) & 0xFF == 1 and (-r) >> 8 == 4 and off == 3
    # Lz4RawDecompressor.java:48-57
    r, off, _ = oracle.decompress_raw("lz4", b"", 10)
    assert (-r) >> 8 == 1
    assert oracle.decompress_raw("lz4", b"\x00", 0)[0] == 0


def test_snappy_known_answer_errors(oracle):
    # T/snappy/TestSnappyJava.java:52-59 -> "Malformed input: offset=2"
    r, off, _ = oracle.decompress_raw("snappy", bytes([16, 1, 0, 1, 0, 1, 0, 1, 0]), 64)
    assert (-r) & 0xFF == 1 and off == 2
    # T/snappy/AbstractTestSnappy.java:31-46 invalid literal length
    data = bytes([128, 8, 252, 255, 255, 255, 127, 0, 0, 0, 0, 0, 0, 0, 0])
    r, off, _ = oracle.decompress_raw("snappy", data, 1024)
    assert (-r) & 0xFF == 1
    # T/snappy/AbstractTestSnappy.java:48-56 negative length
    import ctypes
    off = ctypes.c_int64(0)
    buf = np.frombuffer(bytes([255, 255, 255, 255, 8]), dtype=np.uint8)
    r = oracle.lib.orc_snappy_uncompressed_length(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 5, ctypes.byref(off))
    assert (-r) >> 8 == 9


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_roundtrip_matrix_vs_reference_natives(oracle, refnative, codec, sample_blocks, synthetic_cases):
    """AbstractTestCompression.testDecompress / testCompress with (under test = oracle port,
    verify = the reference's bundled native library)."""
    for blk in synthetic_cases + sample_blocks:
        c = oracle.compress(codec, blk)
        assert len(c) <= oracle.max_compressed_length(codec, len(blk))
        assert refnative.decompress(codec, c, len(blk)) == blk           # verify decompressor accepts our stream
        assert oracle.decompress(codec, c, len(blk)) == blk              # exact-size output buffer
        c2 = refnative.compress(codec, blk)
        assert oracle.decompress(codec, c2, len(blk)) == blk             # we decode the verify compressor's stream
        assert oracle.decompress(codec, c2, len(blk) + 1021) == blk      # padded output


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_small_literal_roundtrip(oracle, refnative, codec):
    # AbstractTestCompression.testRoundTripSmallLiteral :617-648
    data = bytes(range(256))
    for n in range(1, 256):
        c = oracle.compress(codec, data[:n])
        assert oracle.decompress(codec, c, n) == data[:n]
        assert refnative.decompress(codec, c, n) == data[:n]


def test_max_compressed_length(oracle):
    assert oracle.max_compressed_length("lz4", 65536) == 65809
    assert oracle.max_compressed_length("snappy", 65536) == 76490


def test_config0_lz4_whole_file_roundtrip_single_thread(oracle, refnative, pieces):
    """BASELINE.json configs[0] (CPU plumbing, no GPU): the reference calls the codec on whole files
    (AbstractTestCompression.java:362-393).  The LZ4 port compresses and decompresses the whole corpus sample as ONE input,
    single-threaded, the result is byte-identical, and the reference's bundled liblz4 decodes the same stream."""
    import time
    whole = np.concatenate(pieces).tobytes()
    t0 = time.perf_counter()
    c = oracle.compress("lz4", whole)
    t1 = time.perf_counter()
    back = oracle.decompress("lz4", c, len(whole))
    t2 = time.perf_counter()
    assert back == whole
    assert refnative.decompress("lz4", c, len(whole)) == whole
    assert len(c) <= oracle.max_compressed_length("lz4", len(whole))
    mib = len(whole) / (1 << 20)
    print(f"lz4 port, 1 thread, {mib:.1f} MiB: compress {mib / (t1 - t0):.0f} MiB/s, decompress {mib / (t2 - t1):.0f} MiB/s, ratio {len(c) / len(whole):.3f}")


def test_xxh64_hash_long_equals_bytes(oracle):
    # AbstractTestXxHash64.java:321-341: hash(long) == hash of its 8 little-endian bytes, with and without seed
    PRIME32 = 0x9E3779B1
    for v in (0x0102030405060708, 0, 12345, 0x7FFFFFFFFFFFFFFF, 0xDEADBEEFCAFEBABE):
        for seed in (0, PRIME32):
            assert oracle.xxh64_long(v, seed) == oracle.xxh64(v.to_bytes(8, "little"), seed)
