"""CPU tests of the f1 row's checkers: the XXH32 oracle against the reference's known answers and its bundled libxxhash, the
LZ4 frame oracle (oracle/lz4_frame_oracle.py = Lz4FrameCompression.java restated) against an independent implementation of
the frame format (the image's liblz4 LZ4F) in both directions and against the hand-built frames of
T/lz4/TestLz4FrameDecompressor.java:61-230, and that the C ABI exports the XXH32 entry points."""
import struct

import numpy as np
import pytest

from oracle import lz4_frame_oracle as fo
from lz4f_native import Lz4fNative

CONTENT = b"the quick brown fox jumps over the lazy dog"


def build_frame(oracle, flags, bd=0x70, content_size=len(CONTENT), block_checksum=None, content_checksum=None, content=CONTENT):
    """FrameBuilder of the reference test (TestLz4FrameDecompressor.java:268-340): one stored block holding `content`."""
    desc = bytes([fo.FLG_VERSION | flags, bd])
    if flags & fo.FLG_CONTENT_SIZE:
        desc += struct.pack("<q", content_size)
    if flags & fo.FLG_DICTIONARY_ID:
        desc += struct.pack("<I", 0xCAFEBABE)
    f = struct.pack("<I", fo.MAGIC) + desc + bytes([(oracle.xxh32(desc) >> 8) & 0xFF])
    f += struct.pack("<I", len(content) | fo.UNCOMPRESSED_BLOCK_FLAG) + content
    if flags & fo.FLG_BLOCK_CHECKSUM:
        f += struct.pack("<I", block_checksum if block_checksum is not None else oracle.xxh32(content))
    f += struct.pack("<I", 0)
    if flags & fo.FLG_CONTENT_CHECKSUM:
        f += struct.pack("<I", content_checksum if content_checksum is not None else oracle.xxh32(content))
    return f


def skippable(content):
    return struct.pack("<II", fo.SKIPPABLE_MAGIC, len(content)) + content


def reference_cases(oracle):
    """(frame, output capacity, expected bytes or expected message fragment) of T/lz4/TestLz4FrameDecompressor.java:61-230."""
    I = fo.FLG_BLOCK_INDEPENDENCE
    x = oracle.xxh32(CONTENT)
    ok = build_frame(oracle, I)
    bad_magic = bytearray(ok); bad_magic[0] ^= 0xFF
    bad_version = bytearray(ok); bad_version[4] &= 0x3F
    bad_hc = bytearray(ok); bad_hc[6] ^= 0xFF
    cc = build_frame(oracle, I | fo.FLG_CONTENT_CHECKSUM)
    sk = skippable(b"ignored metadata")
    n = len(CONTENT)
    return [
        (cc, n, CONTENT),
        (build_frame(oracle, I | fo.FLG_CONTENT_CHECKSUM, content_checksum=x ^ 1), n, "invalid content checksum"),
        (build_frame(oracle, I | fo.FLG_BLOCK_CHECKSUM), n, CONTENT),
        (build_frame(oracle, I | fo.FLG_BLOCK_CHECKSUM, block_checksum=x ^ 1), n, "invalid block checksum"),
        (build_frame(oracle, I | fo.FLG_CONTENT_SIZE), n, CONTENT),
        (build_frame(oracle, I | fo.FLG_CONTENT_SIZE, content_size=n + 1), n, "content size does not match"),
        (build_frame(oracle, 0), n, "linked blocks are not supported"),
        (build_frame(oracle, I | fo.FLG_DICTIONARY_ID), n, "dictionary are not supported"),
        (build_frame(oracle, I | fo.FLG_RESERVED_MASK), n, "reserved bits"),
        (build_frame(oracle, I, bd=0x71), n, "reserved bits"),
        (bytes(bad_magic), n, "magic number"),
        (bytes(bad_version), n, "Unsupported LZ4 frame version"),
        (build_frame(oracle, I, bd=0x10), n, "block maximum size"),
        (bytes(bad_hc), n, "invalid header checksum"),
        (cc + cc + cc, 3 * n, CONTENT * 3),
        (sk + ok + sk + ok + sk, 2 * n, CONTENT * 2),
        (skippable(b"") + ok, n, CONTENT),
        (ok + sk[:-1], n, "Truncated LZ4 skippable frame"),
        (ok + bytes([1, 2, 3, 4, 5]), n, "magic number"),
        (ok, n - 1, "Output buffer too small"),
        (ok[:5], n, "Input is too short"),
    ]


def test_xxh32_oracle_known_answers_and_native_parity(oracle, refnative):
    assert oracle.xxh32(b"") == 0x02CC5D05 and oracle.xxh32(b"abc") == 0x32D153FF      # T/xxhash/TestXxHash32.java:45-46
    rng = np.random.default_rng(3)
    data = bytes(rng.integers(0, 256, 4099, dtype=np.uint8))
    for seed in (0, 1, 0x9E3779B1, 0xFFFFFFFF, 0x7FFFFFFF, 0x80000000):                 # SEEDS :30
        for length in list(range(0, 67)) + [127, 128, 129, 1023, 1024, 4099]:
            assert oracle.xxh32(data[:length], seed) == refnative.xxh32(data[:length], seed), (seed, length)
    assert oracle.xxh32(data[5:1000]) == refnative.xxh32(data[5:1000])                  # any alignment


def test_frame_oracle_on_the_reference_cases(oracle):
    for i, (frame, cap, want) in enumerate(reference_cases(oracle)):
        if isinstance(want, bytes):
            assert fo.decompress(oracle, frame, cap) == want, i
        else:
            with pytest.raises(fo.FrameError, match=want):
                fo.decompress(oracle, frame, cap)


def test_frame_oracle_interoperates_with_liblz4(oracle, pieces):
    nat = Lz4fNative()
    rng = np.random.default_rng(5)
    blobs = [b"", b"a", CONTENT * 100, pieces[0][:300000].tobytes(), bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
             np.concatenate(pieces[3:6])[:4200000].tobytes() + CONTENT]              # more than one 4 MiB block
    for k, blob in enumerate(blobs):
        f = fo.compress(oracle, blob)
        assert fo.decompress(oracle, f, len(blob)) == blob
        assert nat.decompress(f, len(blob)) == blob                                   # liblz4 reads what the reference rules write
        for bs, bsum, csum, csize in ((4, False, False, False), (5, True, False, True), (6, False, True, False), (7, True, True, True)):
            g = nat.compress(blob, bs, bsum, csum, csize)
            assert fo.decompress(oracle, g, len(blob)) == blob, (k, bs)              # and the reference rules read liblz4's frames


def test_c_abi_exports_xxh32():
    from aircompressor_b200 import _native as N
    assert {"acc_xxh32", "acc_xxh32_batch"} <= set(N.exported_symbols())


class _MockEngine:
    """Stands in for BatchEngine in the HOST-LOGIC test below: same run_host contract, blocks decoded / compressed by the oracle."""

    def __init__(self, oracle):
        self.o = oracle

    def run_host(self, op, src, so, sl, dst, do, dc):
        import aircompressor_b200 as acb
        n = len(so)
        out_len, status = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int32)
        for i in range(n):
            blk = bytes(src[int(so[i]):int(so[i]) + int(sl[i])])
            if op == acb.OP_LZ4_DECOMPRESS:
                r, off, data = self.o.decompress_raw("lz4", blk, int(dc[i]))
                if r < 0:
                    status[i], out_len[i] = -r, off
                else:
                    dst[int(do[i]):int(do[i]) + r] = data[:r]
                    out_len[i] = r
            else:
                c = self.o.compress("lz4", blk)
                dst[int(do[i]):int(do[i]) + len(c)] = np.frombuffer(c, dtype=np.uint8)
                out_len[i] = len(c)
        return out_len, status


class _MockXxh32:
    def __init__(self, oracle):
        self.o = oracle

    def hash(self, arr, offset, length, seed=0):
        return self.o.xxh32(bytes(arr[offset:offset + length]), seed)

    def hash_many(self, arr, offsets, lengths, seed=0):
        return np.array([self.o.xxh32(bytes(arr[o:o + n]), seed) for o, n in zip(offsets, lengths)], dtype=np.int64)


def test_frame_host_logic_reports_in_the_order_of_the_java_loop(oracle, pieces):
    """The frame decoder's host side (header walk first, blocks as a batch, errors in the order the sequential Java loop meets
    them) with the GPU calls replaced by the oracle: every outcome -- bytes, or message and offset -- equals the frame oracle's.
    (The product has no such fallback; tests/test_gpu_lz4_frame.py runs the same cases through the CUDA library.)"""
    import aircompressor_b200 as acb
    from aircompressor_b200.lz4_frame import Lz4FrameCudaDecompressor
    dec = object.__new__(Lz4FrameCudaDecompressor)
    dec._engine, dec._x = _MockEngine(oracle), _MockXxh32(oracle)
    nat = Lz4fNative()
    rng = np.random.default_rng(17)
    blob = pieces[1][:200000].tobytes()
    cases = [(f, cap) for f, cap, _w in reference_cases(oracle)]
    base = [fo.compress(oracle, blob), nat.compress(blob, 4, True, True, True), nat.compress(blob, 5, True, False, False)]
    cases += [(f, len(blob)) for f in base] + [(base[0] + base[1], 2 * len(blob))]
    for f in base:
        for _ in range(40):
            m = bytearray(f)
            kind = rng.integers(0, 3)
            if kind == 0:
                m = m[:rng.integers(1, len(m))]
            elif kind == 1:
                m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)
            else:
                m[rng.integers(0, min(len(m), 24))] = rng.integers(0, 256)
            cases.append((bytes(m), len(blob) if rng.integers(0, 4) else int(rng.integers(0, len(blob)))))
    n_bad = 0
    for i, (frame, cap) in enumerate(cases):
        out = np.full(cap + 8, 0xA5, dtype=np.uint8)
        try:
            want = fo.decompress(oracle, frame, cap)
        except fo.FrameError as e:
            n_bad += 1
            with pytest.raises(acb.MalformedInputException) as got:
                dec.decompress(frame, 0, len(frame), out, 0, cap)
            assert (got.value.reason, got.value.offset) == (e.reason, e.offset), i
            continue
        except fo.BlockError:
            n_bad += 1
            with pytest.raises(acb.MalformedInputException):
                dec.decompress(frame, 0, len(frame), out, 0, cap)
            continue
        n = dec.decompress(frame, 0, len(frame), out, 0, cap)
        assert n == len(want) and out[:n].tobytes() == want and (out[cap:] == 0xA5).all(), i
    assert n_bad > 40


def test_frame_host_logic_splits_large_calls_into_several_batches(oracle, pieces):
    """A call whose blocks need more scratch than the per-batch budget is decoded in several batches (host logic, mock engine)."""
    from aircompressor_b200.lz4_frame import Lz4FrameCudaDecompressor
    dec = object.__new__(Lz4FrameCudaDecompressor)
    dec._engine, dec._x = _MockEngine(oracle), _MockXxh32(oracle)
    dec.BLOCKS_PER_BATCH_BYTES = 200000                                  # three 64 KiB slots per batch
    blob = (CONTENT * 3 + bytes(range(256))) * 1800                       # 693,000 compressible bytes
    frame = Lz4fNative().compress(blob, 4, True, True, True)              # eleven 64 KiB blocks, block + content checksums
    calls = []
    run = dec._engine.run_host
    dec._engine.run_host = lambda *a: (calls.append(len(a[2])), run(*a))[1]
    out = np.zeros(len(blob), dtype=np.uint8)
    assert dec.decompress(frame, 0, len(frame), out, 0, len(out)) == len(blob) and out.tobytes() == blob
    assert len(calls) >= 3 and sum(calls) == 11
