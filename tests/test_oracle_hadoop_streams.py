"""CPU tests of the Hadoop block stream adapters' HOST LOGIC (row f2, LZ4 and Snappy block streams): the batched streams of
aircompressor_b200/hadoop_streams.py with the GPU engine replaced by the oracle must write the reference's framing and read
back what the sequential reader of the reference reads -- including blocks of several chunks (Hadoop's own codecs write
those), zero-length blocks, and the error a damaged stream ends with after the bytes in front of it.
(The product has no such fallback; tests/test_gpu_hadoop_streams.py runs the same cases through the CUDA library.)"""
import io
import struct

import numpy as np
import pytest

from oracle import hadoop_stream_oracle as ho
from test_oracle_lz4_frame import _MockEngine


class _MockEngine2(_MockEngine):
    """also compresses Snappy / decompresses Snappy"""

    def run_host(self, op, src, so, sl, dst, do, dc):
        import aircompressor_b200 as acb
        codec = "lz4" if op in (acb.OP_LZ4_COMPRESS, acb.OP_LZ4_DECOMPRESS) else "snappy"
        n = len(so)
        out_len, status = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int32)
        for i in range(n):
            blk = bytes(src[int(so[i]):int(so[i]) + int(sl[i])])
            if op in (acb.OP_LZ4_DECOMPRESS, acb.OP_SNAPPY_DECOMPRESS):
                r, off, data = self.o.decompress_raw(codec, blk, int(dc[i]))
                if r < 0:
                    status[i], out_len[i] = -r, off
                else:
                    dst[int(do[i]):int(do[i]) + r] = data[:r]
                    out_len[i] = r
            else:
                c = self.o.compress(codec, blk)
                dst[int(do[i]):int(do[i]) + len(c)] = np.frombuffer(c, dtype=np.uint8)
                out_len[i] = len(c)
        return out_len, status


def streams_for(codec):
    from aircompressor_b200 import hadoop_streams as hs
    return (hs.Lz4HadoopCudaOutputStream, hs.Lz4HadoopCudaInputStream) if codec == "lz4" else (hs.SnappyHadoopCudaOutputStream, hs.SnappyHadoopCudaInputStream)


def read_all(stream_obj):
    """-> (bytes delivered, exception or None), reading in odd sizes like a real consumer"""
    out = bytearray()
    try:
        for size in [1, 7, 100000] * 100000:
            d = stream_obj.read(size)
            if not d:
                return bytes(out), None
            out += d
    except Exception as e:                      # noqa: BLE001 -- the exception IS the result
        return bytes(out), e


def same_failure(mine, ref):
    if ref is None:
        return mine is None
    if isinstance(ref, tuple):                  # the raw block decoder rejected a chunk
        import aircompressor_b200 as acb
        return isinstance(mine, acb.MalformedInputException)
    return type(mine) is type(ref) and str(mine) == str(ref)


def cases(oracle, codec, pieces, buffer_size):
    rng = np.random.default_rng(29)
    data = [b"", b"a", pieces[0][:3 * buffer_size + 1234].tobytes(), bytes(rng.integers(0, 256, buffer_size + 77, dtype=np.uint8)),
            pieces[2][:buffer_size - ho.overhead(codec, buffer_size)].tobytes()]
    good = [ho.write_stream(oracle, codec, d, buffer_size) for d in data]
    # a block of several chunks (what Hadoop's own codecs write), a zero-length block in front, two streams back to back
    d = pieces[1][:50000].tobytes()
    parts = [d[:20000], d[20000:20001], d[20001:]]
    multi = struct.pack(">I", 0) + struct.pack(">I", len(d)) + b"".join(struct.pack(">I", len(c)) + c for c in (oracle.compress(codec, p) for p in parts))
    good += [multi, multi + good[2], good[2] + multi + good[1]]
    bad = []
    for s in (good[2], multi + good[2]):
        for _ in range(10):
            m = bytearray(s)
            kind = rng.integers(0, 3)
            if kind == 0:
                m = m[:rng.integers(1, len(m))]
            elif kind == 1:
                m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)
            else:
                m[rng.integers(8, 40)] = rng.integers(0, 256)
            bad.append(bytes(m))
    return data, good, bad


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_stream_host_logic_matches_the_reference_streams(oracle, pieces, codec):
    Out, In = streams_for(codec)
    eng = _MockEngine2(oracle)
    bs = 16 * 1024
    data, good, bad = cases(oracle, codec, pieces, bs)
    for d in data:                                                 # writer: the reference's framing, chunk for chunk
        sink = io.BytesIO()
        w = Out(sink, buffer_size=bs, batch_chunks=3, engine=eng)
        for pos in range(0, len(d), 5000):
            w.write(d[pos:pos + 5000])
            if pos == 10000:
                w.flush()
        w.finish()
        mine = sink.getvalue()
        assert mine == ho.write_stream(oracle, codec, d, bs)        # the mock compresses like the reference: byte-identical streams
        assert ho.read_stream(oracle, codec, mine, bs) == (d, None)
    n_bad = 0
    for s in good + bad:
        want, werr = ho.read_stream(oracle, codec, s, bs)
        for batch in (1, 4, 64):
            got, gerr = read_all(In(io.BytesIO(s), buffer_size=bs, batch_chunks=batch, engine=eng))
            assert got == want, (codec, batch, len(got), len(want))
            assert same_failure(gerr, werr), (codec, batch, gerr, werr)
        n_bad += werr is not None
    assert n_bad >= 5
