import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aircompressor_b200 as acb, benchdata
from oracle.pyoracle import Oracle
orc = Oracle(); eng = acb.BatchEngine(0); eng.set_tuning(1, 2)
blocks = benchdata.cut_blocks(benchdata.load_pieces()[1], 65536)
streams = [np.frombuffer(orc.compress("lz4", b.tobytes()), dtype=np.uint8) for b in blocks]
src, so, sl = benchdata.pack(streams); raw, ro, rl = benchdata.pack(blocks)
back = np.zeros_like(raw)
dlen, st = eng.run_host(acb.OP_LZ4_DECOMPRESS, src, so, sl, back, ro, rl)
L = acb.lib(); L.acc_debug_tpb_stats.argtypes = [C.c_void_p]
stats = (C.c_ulonglong * 4)(); L.acc_debug_tpb_stats(stats)
print("ok", bool((st == 0).all() and np.array_equal(back, raw)), "blocks", stats[0], "fallbacks", stats[1], flush=True)
os._exit(0)
