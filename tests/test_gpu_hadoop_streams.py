"""GPU test of the Hadoop block stream adapters (row f2, LZ4 and Snappy block streams) through the CUDA library: streams written
by the batched GPU writer carry the reference's framing (chunk boundaries and length words; the payload bytes are this
library's compressor's) and are read back by the reference's sequential reader rules (oracle/hadoop_stream_oracle.py); streams
written by the reference rules -- including blocks of several chunks and damaged streams -- are read by the batched GPU reader
with the same bytes and the same final error."""
import io

import pytest

from oracle import hadoop_stream_oracle as ho
from test_oracle_hadoop_streams import cases, read_all, same_failure, streams_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_streams_through_the_gpu(oracle, pieces, codec):
    Out, In = streams_for(codec)
    bs = 64 * 1024
    data, good, bad = cases(oracle, codec, pieces, bs)
    for d in data:
        sink = io.BytesIO()
        w = Out(sink, buffer_size=bs, batch_chunks=3)
        for pos in range(0, len(d), 50000):
            w.write(d[pos:pos + 50000])
        w.finish()
        mine = sink.getvalue()
        ref = ho.write_stream(oracle, codec, d, bs)
        assert [u for u, _c in ho.chunk_lengths(mine)] == [u for u, _c in ho.chunk_lengths(ref)]     # same chunk geometry
        assert ho.read_stream(oracle, codec, mine, bs) == (d, None)                                   # the reference reader reads it
        assert read_all(In(io.BytesIO(mine), buffer_size=bs)) == (d, None)
    n_bad = 0
    for s in good + bad:
        want, werr = ho.read_stream(oracle, codec, s, bs)
        for batch in (1, 64):
            got, gerr = read_all(In(io.BytesIO(s), buffer_size=bs, batch_chunks=batch))
            assert got == want, (codec, batch, len(got), len(want))
            assert same_failure(gerr, werr), (codec, batch, gerr, werr)
        n_bad += werr is not None
    assert n_bad >= 5
