import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aircompressor_b200 as acb, benchdata
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = acb.BatchEngine(0)
pieces = benchdata.load_pieces()[1]
blocks = benchdata.cut_blocks(pieces, 128 * 1024)
blocks = (blocks * (n // len(blocks) + 1))[:n]
src, so, sl = benchdata.pack(blocks)
L = acb.lib()
caps = np.array([L.acc_zstd_compress_bound(int(x)) for x in sl], dtype=np.int64)
do = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
comp = np.zeros(int(caps.sum()), dtype=np.uint8)
clen, st = eng.run_host(acb.OP_ZSTD_COMPRESS, src, so, sl, comp, do, caps)
print("compress ok", (st == 0).all(), clen.sum() / sl.sum())
back = np.zeros_like(src)
dlen, st = eng.run_host(acb.OP_ZSTD_DECOMPRESS, comp, do, clen, back, so, sl)
print("decompress ok", (st == 0).all(), np.array_equal(back, src))
