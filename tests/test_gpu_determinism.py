"""The encoders are deterministic: compressing the same batch twice (different launches, different work-stealing order)
gives byte-identical streams, like the reference's compressors do for the same input."""
import numpy as np
import pytest

import aircompressor_b200 as acb
import benchdata

pytestmark = pytest.mark.gpu
OPS = {"lz4": acb.OP_LZ4_COMPRESS, "snappy": acb.OP_SNAPPY_COMPRESS, "zstd": acb.OP_ZSTD_COMPRESS}


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
def test_compress_twice_gives_identical_streams(engine, pieces, codec):
    blocks = benchdata.cut_blocks(pieces, 128 * 1024 if codec == "zstd" else 64 * 1024)
    blocks = (blocks * 8)[:1024]                 # enough inputs that the CTAs pick them up in a different order every launch
    src, so, sl = benchdata.pack(blocks)
    bound = getattr(acb.lib(), f"acc_{codec}_compress_bound")
    caps = np.array([bound(int(n)) for n in sl], dtype=np.int64)
    do = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    runs = []
    for _ in range(3):
        comp = np.zeros(int(caps.sum()), dtype=np.uint8)
        clen, st = engine.run_host(OPS[codec], src, so, sl, comp, do, caps)
        assert (st == 0).all()
        runs.append((clen, comp))
    for clen, comp in runs[1:]:
        assert np.array_equal(clen, runs[0][0])
        for i in range(len(blocks)):
            assert np.array_equal(comp[do[i]:do[i] + clen[i]], runs[0][1][do[i]:do[i] + clen[i]]), (codec, i)
