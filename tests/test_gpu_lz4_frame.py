"""GPU parity tests of the f1 row (through the C ABI): batched XXH32 against the oracle (= XxHash32JavaHasher rules, pinned to
the reference's known answers and bundled libxxhash), and the LZ4 frame codec (aircompressor_b200/lz4_frame.py: host-side
framing like the reference, blocks and block checksums as GPU batches) against the frame oracle
(oracle/lz4_frame_oracle.py = Lz4FrameCompression.java restated): same bytes, and for malformed input the same message and
offset -- on the reference's own hand-built cases (T/lz4/TestLz4FrameDecompressor.java:61-230), on frames written by the
reference rules, by this codec and by liblz4's LZ4F (block sizes 64 KiB - 4 MiB, block / content checksums, content size),
and on truncated and bit-flipped frames."""
import numpy as np
import pytest

import aircompressor_b200 as acb
from oracle import lz4_frame_oracle as fo
from lz4f_native import Lz4fNative
from test_oracle_lz4_frame import CONTENT, reference_cases

pytestmark = pytest.mark.gpu


def test_xxh32_matches_oracle(engine, oracle):
    rng = np.random.default_rng(9)
    data = rng.integers(0, 256, 300000, dtype=np.uint8)
    lens = list(range(0, 70)) + [127, 128, 129, 255, 1000, 4096, 65535, 65536, 100001]
    offs, pos = [], 0
    for i, n in enumerate(lens):
        pos += (i * 3) % 7                       # every alignment
        offs.append(pos)
        pos += n
    out, st = engine.run_host(acb.OP_XXH32, data, np.array(offs), np.array(lens), None, None, None)
    assert (st == 0).all()
    for i, n in enumerate(lens):
        assert int(out[i]) == oracle.xxh32(data[offs[i]:offs[i] + n].tobytes()), (i, n)
    h = acb.XxHash32CudaHasher()
    assert h.hash(b"") == 0x02CC5D05 and h.hash(b"abc") == 0x32D153FF                    # T/xxhash/TestXxHash32.java:45-46
    for seed in (0, 1, 0x9E3779B1, -1, 0x7FFFFFFF, -0x80000000):                           # SEEDS :30
        for n in (0, 3, 16, 17, 1000):
            assert h.hash(data[:n], 0, n, seed) & 0xFFFFFFFF == oracle.xxh32(data[:n].tobytes(), seed & 0xFFFFFFFF), (seed, n)
    assert h.hash(data, 5, 995) & 0xFFFFFFFF == oracle.xxh32(data[5:1000].tobytes())
    many = h.hash_many(data, [0, 1, 2, 3], [100, 100, 100, 100])
    assert [int(x) for x in many] == [oracle.xxh32(data[k:k + 100].tobytes()) for k in range(4)]


def _same_outcome(dec, oracle, frame, cap):
    out = np.full(cap + 64, 0xA5, dtype=np.uint8)
    try:
        want = fo.decompress(oracle, frame, cap)
    except fo.FrameError as e:
        with pytest.raises(acb.MalformedInputException) as got:
            dec.decompress(frame, 0, len(frame), out, 0, cap)
        assert got.value.reason == e.reason and got.value.offset == e.offset, (got.value.reason, got.value.offset, e.reason, e.offset)
        return None
    except fo.BlockError:
        with pytest.raises(acb.MalformedInputException):      # the raw block decoder's own report
            dec.decompress(frame, 0, len(frame), out, 0, cap)
        return None
    n = dec.decompress(frame, 0, len(frame), out, 0, cap)
    assert n == len(want) and out[:n].tobytes() == want
    assert (out[cap:] == 0xA5).all()
    return want


def test_reference_frame_cases(oracle):
    dec = acb.Lz4FrameCudaDecompressor()
    for i, (frame, cap, want) in enumerate(reference_cases(oracle)):
        got = _same_outcome(dec, oracle, frame, cap)
        if isinstance(want, bytes):
            assert got == want, i
        else:
            assert got is None, i


def test_frames_round_trip_and_interoperate(oracle, pieces):
    comp, dec, nat = acb.Lz4FrameCudaCompressor(), acb.Lz4FrameCudaDecompressor(), Lz4fNative()
    rng = np.random.default_rng(13)
    blobs = [b"", b"a", CONTENT * 100, pieces[0][:300000].tobytes(), bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
             np.concatenate(pieces[3:7])[:9000000].tobytes() + CONTENT]                  # three 4 MiB blocks
    frames = []
    for k, blob in enumerate(blobs):
        cap = comp.maxCompressedLength(len(blob))
        assert cap == 7 + 4 + len(blob) + 4 * ((len(blob) + (4 << 20) - 1) // (4 << 20))   # Lz4FrameCompression.java:67-80
        buf = np.zeros(cap + 9, dtype=np.uint8)
        n = comp.compress(blob, 0, len(blob), buf, 5, cap)
        mine = buf[5:5 + n].tobytes()
        assert mine[:7] == fo.compress(oracle, b"")[:7]                                  # the header the reference writes
        assert fo.decompress(oracle, mine, len(blob)) == blob                            # the reference rules read this codec's frames
        assert nat.decompress(mine, len(blob)) == blob                                   # and so does liblz4
        frames.append((mine, blob))
        frames.append((fo.compress(oracle, blob), blob))                                 # frames of the reference rules
        for bs, bsum, csum, csize in ((4, False, False, False), (5, True, False, True), (6, False, True, False), (7, True, True, True)):
            frames.append((nat.compress(blob, bs, bsum, csum, csize), blob))
        if len(blob) > 100:
            with pytest.raises(acb.IllegalArgumentException, match="Output buffer too small"):
                comp.compress(blob, 0, len(blob), buf, 0, 50)
    for f, blob in frames:
        assert _same_outcome(dec, oracle, f, len(blob) + 11) == blob
    # several frames and skippable frames in one call = one batch of blocks
    cat = frames[3][0] + b"\x50\x2a\x4d\x18\x03\x00\x00\x00abc" + frames[20][0] + frames[9][0]
    want = frames[3][1] + frames[20][1] + frames[9][1]
    assert _same_outcome(dec, oracle, cat, len(want)) == want


def test_malformed_frames_report_like_the_reference(oracle, pieces):
    dec, nat = acb.Lz4FrameCudaDecompressor(), Lz4fNative()
    rng = np.random.default_rng(17)
    blob = pieces[1][:200000].tobytes()
    base = [fo.compress(oracle, blob), nat.compress(blob, 4, True, True, True), nat.compress(blob, 5, True, False, False)]
    n_bad = 0
    for f in base:
        for _ in range(12):
            m = bytearray(f)
            kind = rng.integers(0, 3)
            if kind == 0:
                m = m[:rng.integers(1, len(m))]
            elif kind == 1:
                m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)
            else:
                m[rng.integers(0, min(len(m), 24))] = rng.integers(0, 256)
            cap = len(blob) if rng.integers(0, 4) else int(rng.integers(0, len(blob)))
            if _same_outcome(dec, oracle, bytes(m), cap) is None:
                n_bad += 1
    assert n_bad > 15
