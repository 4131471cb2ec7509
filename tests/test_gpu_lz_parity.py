"""GPU parity tests (through the C ABI): LZ4 and Snappy block codecs vs the oracle.

Decode must be bit-exact with the Java decoder restatement: same bytes, same length, and for corrupt
input the same status word and error offset.  Encode must produce streams the oracle decoder (= Java
decoder rules) and the reference's native library both round-trip (AbstractTestCompression.java:74-99,
:362-393).
"""
import numpy as np
import pytest

import aircompressor_b200 as acb
import benchdata

pytestmark = pytest.mark.gpu
OPS = {"lz4": (acb.OP_LZ4_COMPRESS, acb.OP_LZ4_DECOMPRESS), "snappy": (acb.OP_SNAPPY_COMPRESS, acb.OP_SNAPPY_DECOMPRESS)}


def _pack(chunks, pad=0):
    lens = np.array([len(c) for c in chunks], dtype=np.int64)
    offs = np.zeros(len(chunks), dtype=np.int64)
    pos = 0
    for i, c in enumerate(chunks):
        offs[i] = pos
        pos += len(c) + pad
    buf = np.zeros(max(pos, 1), dtype=np.uint8)
    for i, c in enumerate(chunks):
        buf[offs[i]:offs[i] + len(c)] = np.frombuffer(c, dtype=np.uint8)
    return buf, offs, lens


def _gpu_decompress(engine, codec, streams, caps, guard=0):
    src, so, sl = _pack(streams, pad=3)  # odd padding: unaligned block starts
    caps = np.asarray(caps, dtype=np.int64)
    do = np.zeros(len(streams), dtype=np.int64)
    pos = 0
    for i, c in enumerate(caps):
        do[i] = pos
        pos += int(c) + guard
    dst = np.full(max(pos, 1), 0xA5, dtype=np.uint8)
    out_len, status = engine.run_host(OPS[codec][1], src, so, sl, dst, do, caps)
    return dst, do, out_len, status


@pytest.fixture(params=[2, 1], ids=["record-path", "step-decoder"])
def decoder(request, engine):
    """the decode tests run against the record path (parse + execute kernels, lz_records.cuh: Snappy's default) and against the
    step decoder alone (LZ4's default, and what every block's tail resumes in); acc_set_tuning key 1"""
    engine.set_tuning(1, request.param)
    yield request.param
    engine.set_tuning(1, 0)


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_decompress_matches_oracle_on_corpus(engine, oracle, refnative, codec, sample_blocks, synthetic_cases, decoder):
    blocks = [b for b in synthetic_cases + sample_blocks]
    streams, caps, want = [], [], []
    for i, blk in enumerate(blocks):
        for comp in (oracle, refnative):
            c = comp.compress(codec, blk)
            streams.append(c)
            caps.append(len(blk) + (1021 if i % 3 == 0 else 0))   # exact-size and padded outputs
            want.append(blk)
    dst, do, out_len, status = _gpu_decompress(engine, codec, streams, caps, guard=100)
    for i, blk in enumerate(want):
        r, off, ref_out = oracle.decompress_raw(codec, streams[i], caps[i])
        assert status[i] == 0 and r == len(blk), (i, status[i], out_len[i], r)
        assert out_len[i] == len(blk)
        got = dst[do[i]:do[i] + len(blk)].tobytes()
        assert got == blk and got == ref_out[:r].tobytes()
        # AbstractTestCompression.testDecompressionBufferOverrun :131-163 -- nothing past maxOutputLength is touched
        assert (dst[do[i] + caps[i]:do[i] + caps[i] + 100] == 0xA5).all()


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_decompress_error_parity_on_corrupt_streams(engine, oracle, codec, sample_blocks, decoder):
    """Bit flips, truncations and wrong capacities: GPU status word and offset == oracle's."""
    rng = np.random.default_rng(99)
    streams, caps = [], []
    base = [b for b in sample_blocks if 0 < len(b) <= 65536][:24]
    for blk in base:
        c = bytearray(oracle.compress(codec, blk))
        for _ in range(6):
            m = bytearray(c)
            kind = rng.integers(0, 4)
            if kind == 0 and len(m) > 4:
                m = m[:rng.integers(1, len(m))]                         # truncated
            elif kind == 1:
                for _k in range(rng.integers(1, 4)):
                    m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)  # bit flips
            elif kind == 2:
                m[rng.integers(0, min(len(m), 64))] = rng.integers(0, 256)
            streams.append(bytes(m))
            caps.append(len(blk) if kind != 3 else int(rng.integers(0, len(blk))))  # too-small output
    # known-answer vectors of the reference tests
    if codec == "lz4":
        streams += [bytes([15, 0, 0, 255, 255, 138, 49, 255, 255, 0]), b"", b"\x00", b"\x00", b"\x10"]
        caps += [1024, 16, 0, 5, 0]
    else:
        streams += [bytes([16, 1, 0, 1, 0, 1, 0, 1, 0]), bytes([128, 8, 252, 255, 255, 255, 127, 0, 0, 0, 0, 0, 0, 0, 0]),
                    bytes([255, 255, 255, 255, 8]), b"", bytes([0x80])]
        caps += [64, 1024, 64, 8, 8]
    dst, do, out_len, status = _gpu_decompress(engine, codec, streams, caps, guard=64)
    n_bad = 0
    for i, s in enumerate(streams):
        r, off, ref_out = oracle.decompress_raw(codec, s, caps[i])
        if r >= 0:
            assert status[i] == 0 and out_len[i] == r, (i, status[i], out_len[i], r)
            assert dst[do[i]:do[i] + r].tobytes() == ref_out[:r].tobytes()
        else:
            n_bad += 1
            assert status[i] == -r, (i, hex(status[i]), hex(-r), off, out_len[i])
            assert out_len[i] == off, (i, out_len[i], off)
        assert (dst[do[i] + caps[i]:do[i] + caps[i] + 64] == 0xA5).all()
    assert n_bad > 10


def test_lz4_overflow_streams(engine, oracle, decoder):
    # T/lz4/AbstractTestLz4.java:28-67 (shortened: the 9 MB run is what overflows the Java int)
    n = (2**31 - 1) // 255 + 1
    lit = bytes([0xF0]) + b"\xff" * n + bytes([1]) + bytes(20)
    mat = bytes([0x0F, 0, 0]) + b"\xff" * n + bytes([1]) + bytes(10)
    dst, do, out_len, status = _gpu_decompress(engine, "lz4", [lit, mat], [2048, 2048])
    for i, s in enumerate([lit, mat]):
        r, off, _ = oracle.decompress_raw("lz4", s, 2048)
        assert r < 0 and status[i] == -r and out_len[i] == off


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_compress_roundtrips_through_reference_decoders(engine, oracle, refnative, codec, sample_blocks, synthetic_cases):
    blocks = synthetic_cases + sample_blocks + [bytes(range(256))[:n] for n in range(1, 256)]  # + testRoundTripSmallLiteral
    src, so, sl = _pack(blocks, pad=1)
    L = acb.lib()
    bound = getattr(L, f"acc_{codec}_compress_bound")
    caps = np.array([bound(len(b)) for b in blocks], dtype=np.int64)
    do = np.concatenate([[0], np.cumsum(caps + 32)[:-1]]).astype(np.int64)
    dst = np.full(int((caps + 32).sum()), 0x5A, dtype=np.uint8)
    out_len, status = engine.run_host(OPS[codec][0], src, so, sl, dst, do, caps)
    total_in = total_out = 0
    for i, blk in enumerate(blocks):
        assert status[i] == 0, (i, hex(status[i]))
        c = dst[do[i]:do[i] + out_len[i]].tobytes()
        assert 0 < len(c) <= caps[i]
        assert (dst[do[i] + caps[i]:do[i] + caps[i] + 32] == 0x5A).all()
        assert oracle.decompress(codec, c, len(blk)) == blk            # Java decoder rules, exact-size output
        assert refnative.decompress(codec, c, len(blk)) == blk         # independent verify decompressor
        total_in += len(blk)
        total_out += len(c)
    # too-small output -> argument error, like Lz4RawCompressor.java:87-89 / SnappyRawCompressor.java:87-90
    out_len, status = engine.run_host(OPS[codec][0], src, so[5:6], sl[5:6], dst, do[:1], caps[5:6] - 1)
    assert status[0] & 0xFF == 3
    print(f"{codec}: gpu ratio {total_out / total_in:.4f}")


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_gpu_roundtrip_full_batch_property(engine, codec, pieces, decoder):
    """Size-independent property at bench scale: compress -> decompress of every 64 KiB block of the
    corpus sample tiled to 4096 blocks returns the input (checked by comparing whole buffers)."""
    blocks = benchdata.cut_blocks(pieces, 64 * 1024)
    reps = (4096 + len(blocks) - 1) // len(blocks)
    blocks = (blocks * reps)[:4096]
    src, so, sl = benchdata.pack(blocks)
    L = acb.lib()
    bound = getattr(L, f"acc_{codec}_compress_bound")
    caps = np.array([bound(int(n)) for n in sl], dtype=np.int64)
    do = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    comp = np.zeros(int(caps.sum()), dtype=np.uint8)
    clen, st = engine.run_host(OPS[codec][0], src, so, sl, comp, do, caps)
    assert (st == 0).all()
    back = np.zeros_like(src)
    dlen, st = engine.run_host(OPS[codec][1], comp, do, clen, back, so, sl)
    assert (st == 0).all() and (dlen == sl).all()
    assert np.array_equal(back, src)


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_pipelined_host_path_equals_single_pass(engine, oracle, codec, pieces):
    """acc_batch with host pointers cuts large batches into overlapped upload/kernel/download runs (tuning key 3).
    The split must not change a byte: lengths, statuses (some blocks are corrupted on purpose), outputs and the
    guard bytes between output windows are compared between 1 run and 5 runs, for both directions."""
    blocks = benchdata.cut_blocks(pieces, 16 * 1024)[:333]
    src, so, sl = benchdata.pack(blocks)
    bound = getattr(acb.lib(), f"acc_{codec}_compress_bound")
    caps = np.array([bound(int(n)) + 7 for n in sl], dtype=np.int64)
    do = np.concatenate([[0], np.cumsum(caps + 5)[:-1]]).astype(np.int64)     # 5 guard bytes after every window
    results = []
    for chunks in (1, 5):
        engine.set_tuning(3, chunks)
        try:
            comp = np.full(int(do[-1] + caps[-1] + 5), 0x5A, dtype=np.uint8)
            clen, st = engine.run_host(OPS[codec][0], src, so, sl, comp, do, caps)
            assert (st == 0).all()
            bad = comp.copy()
            for i in range(0, len(blocks), 17):       # corrupt every 17th stream: statuses must survive the split too
                bad[do[i] + clen[i] // 2] ^= 0xFF
                bad[do[i] + clen[i] // 2 + 1] ^= 0x81
            back = np.full(len(src) + 64, 0xC3, dtype=np.uint8)
            dlen, dst = engine.run_host(OPS[codec][1], bad, do, clen, back, so, sl)
            results.append((comp, clen, back, dlen, dst))
        finally:
            engine.set_tuning(3, 0)
    a, b = results
    for k in (1, 3, 4):                       # clen, dlen (error offsets included), statuses
        assert np.array_equal(a[k], b[k])
    for i in range(len(blocks)):              # produced bytes; the rest of a window is unspecified (aircompress_cuda.h)
        assert np.array_equal(a[0][do[i]:do[i] + a[1][i]], b[0][do[i]:do[i] + b[1][i]])
        if a[4][i] == 0:
            assert np.array_equal(a[2][so[i]:so[i] + a[3][i]], b[2][so[i]:so[i] + b[3][i]])
    comp, clen, back, dlen, dst = b
    for i in range(len(blocks)):
        if i % 17:
            assert dst[i] == 0 and dlen[i] == sl[i] and np.array_equal(back[so[i]:so[i] + sl[i]], blocks[i])
    assert (dst[::17] != 0).sum() > 0
    assert (back[len(src):] == 0xC3).all()
    for i in range(len(blocks)):
        assert (comp[do[i] + caps[i]:do[i] + caps[i] + 5] == 0x5A).all()


def test_java_shaped_single_block_api():
    c, d = acb.Lz4CudaCompressor(), acb.Lz4CudaDecompressor()
    data = b"XXXXabcdefgh abcdefgh abcdefgh abcdefgh abcdefgh abcdefgh ABC" * 50
    out = bytearray(c.maxCompressedLength(len(data)) + 10)
    n = c.compress(data, 0, len(data), out, 5, len(out) - 5)
    back = bytearray(len(data) + 7)
    m = d.decompress(out, 5, n, back, 7, len(data))
    assert m == len(data) and bytes(back[7:]) == data
    with pytest.raises(acb.IllegalArgumentException, match="Invalid offset or length"):
        d.decompress(out, 5, len(out), back, 0, len(back))
    with pytest.raises(acb.MalformedInputException, match="offset outside destination buffer: offset=3"):
        d.decompress(bytes([15, 0, 0, 255, 255, 138, 49, 255, 255, 0]), 0, 10, bytearray(1024), 0, 1024)
    assert d.decompress(b"\x10", 0, 1, bytearray(0), 0, 0) == -1
    sc, sd = acb.SnappyCudaCompressor(), acb.SnappyCudaDecompressor()
    out = bytearray(sc.maxCompressedLength(len(data)))
    n = sc.compress(data, 0, len(data), out, 0, len(out))
    assert sd.getUncompressedLength(out, 0) == len(data)
    back = bytearray(len(data))
    assert sd.decompress(out, 0, n, back, 0, len(back)) == len(data) and bytes(back) == data
    with pytest.raises(acb.MalformedInputException, match="Malformed input: offset=2"):
        sd.decompress(bytes([16, 1, 0, 1, 0, 1, 0, 1, 0]), 0, 9, bytearray(64), 0, 64)
    # memory-segment style overloads
    seg_out = np.zeros(c.maxCompressedLength(len(data)), dtype=np.uint8)
    n = c.compress(np.frombuffer(data, dtype=np.uint8), seg_out)
    seg_back = np.zeros(len(data), dtype=np.uint8)
    assert d.decompress(seg_out[:n].copy(), seg_back) == len(data) and seg_back.tobytes() == data
