import sys, collections
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import aircompressor_b200 as acb, benchdata
from oracle.pyoracle import Oracle, RefNative
import test_gpu_zstd as T
orc, ref = Oracle(), RefNative()
eng = acb.BatchEngine(0)
pieces = benchdata.load_pieces(sample=True)[1]
out = []
for size, stride in ((64 * 1024, 5), (4 * 1024, 97), (128 * 1024, 7), (1000, 411)):
    blocks = benchdata.cut_blocks(pieces, size)
    out += [blocks[i].tobytes() for i in range(0, len(blocks), stride)]
rng = np.random.default_rng(5)
streams, caps = [], []
base = [b for b in out if 1000 <= len(b) <= 131072][:40]
for blk in base:
    for c in (bytearray(orc.compress("zstd", blk)), bytearray(ref.compress("zstd", blk, 3))):
        for _ in range(8):
            m = bytearray(c); kind = rng.integers(0, 4)
            if kind == 0: m = m[:rng.integers(1, len(m))]
            elif kind == 1:
                for _k in range(rng.integers(1, 3)): m[rng.integers(0, len(m))] ^= 1 << rng.integers(0, 8)
            elif kind == 2: m[rng.integers(0, min(len(m), 48))] = rng.integers(0, 256)
            streams.append(bytes(m)); caps.append(len(blk) if kind != 3 else int(rng.integers(0, len(blk))))
dst, do, out_len, status = T._gpu_decompress(eng, streams, caps)
tab = collections.Counter(); offdiff = collections.Counter()
for i, s in enumerate(streams):
    r, off, _ = orc.decompress_raw("zstd", s, caps[i])
    if r < 0:
        tab[(hex(-r), hex(int(status[i])))] += 1
        if int(status[i]) == -r and out_len[i] != off: offdiff[(hex(-r), int(out_len[i]) - off)] += 1
for k, v in sorted(tab.items()): print(k, v)
print("offset diffs where reason equal:", dict(offdiff))
