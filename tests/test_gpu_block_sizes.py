"""GPU parity across the block-size sweep of BASELINE.json configs[4] (4 KiB - 1 MiB), all three codecs, through the C ABI:
oracle-compressed (= reference algorithm) streams must decode bit-exactly on the GPU, and GPU-compressed streams must decode
with the oracle decoder (= Java decoder rules) to the original bytes."""
import numpy as np
import pytest

import aircompressor_b200 as acb
import benchdata

pytestmark = pytest.mark.gpu
OPS = {"lz4": (acb.OP_LZ4_COMPRESS, acb.OP_LZ4_DECOMPRESS), "snappy": (acb.OP_SNAPPY_COMPRESS, acb.OP_SNAPPY_DECOMPRESS),
       "zstd": (acb.OP_ZSTD_COMPRESS, acb.OP_ZSTD_DECOMPRESS)}


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
@pytest.mark.parametrize("block_kib", [4, 16, 256, 1024])
def test_block_size_sweep(engine, oracle, pieces, codec, block_kib):
    blocks = benchdata.cut_blocks(pieces, block_kib * 1024)
    stride = max(1, len(blocks) // 24)
    blocks = blocks[::stride][:24] + [blocks[-1]]          # includes the ragged final block
    src, so, sl = benchdata.pack(blocks)
    # decode: streams from the oracle compressor
    streams = [np.frombuffer(oracle.compress(codec, b.tobytes()), dtype=np.uint8) for b in blocks]
    cs, co, cl = benchdata.pack(streams)
    back = np.full(len(src) + 16, 0x3C, dtype=np.uint8)
    dlen, st = engine.run_host(OPS[codec][1], cs, co, cl, back, so, sl)
    assert (st == 0).all() and (dlen == sl).all()
    assert np.array_equal(back[:len(src)], src) and (back[len(src):] == 0x3C).all()
    # encode: GPU streams through the oracle decoder
    bound = getattr(acb.lib(), f"acc_{codec}_compress_bound")
    caps = np.array([bound(int(n)) for n in sl], dtype=np.int64)
    do = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    comp = np.zeros(int(caps.sum()), dtype=np.uint8)
    clen, st = engine.run_host(OPS[codec][0], src, so, sl, comp, do, caps)
    assert (st == 0).all() and (clen <= caps).all()
    for i, b in enumerate(blocks):
        assert oracle.decompress(codec, comp[do[i]:do[i] + clen[i]].tobytes(), b.size) == b.tobytes(), (codec, block_kib, i)
