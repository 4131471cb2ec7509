import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aircompressor_b200 as acb
import bench
from oracle.pyoracle import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
orc = Oracle(); eng = acb.BatchEngine(0); dev = torch.device("cuda", 0)
eng.set_tuning(1, 3)
wl = bench.build_workload("lz4", 64, n, orc, bench.host_threads())
def tiled(packed, off, ln):
    reps, stride, stride_al, offs, lens = bench.tile_index(off, ln, n)
    one = torch.zeros(stride_al, dtype=torch.uint8, device=dev); one[:stride] = torch.from_numpy(packed).to(dev)
    return one.repeat(reps), torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev), offs, lens
raw_d, ro_d, rl_d, ro, rl = tiled(wl["raw"], wl["raw_off"], wl["raw_len"])
src_d, so_d, sl_d, so, sl = tiled(wl["comp"], wl["comp_off"], wl["comp_len"])
dst_d = torch.zeros_like(raw_d); ol = torch.zeros(n, dtype=torch.int64, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
L = acb.lib(); L.acc_debug_lz4v3_stats.argtypes = [C.c_void_p]
stats = (C.c_ulonglong * 16)()
for it in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_device(1, src_d.data_ptr(), so_d.data_ptr(), sl_d.data_ptr(), dst_d.data_ptr(), ro_d.data_ptr(), rl_d.data_ptr(), ol.data_ptr(), st.data_ptr(), n, s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    L.acc_debug_lz4v3_stats(stats)
    v = list(stats)
    names = ["blocks", "fallbacks", "parse", "fill", "lit", "match", "flush", "prounds", "mrounds", "total", "seqs"]
    nb = max(v[0], 1)
    print("ms", e0.elapsed_time(e1), {k: round(x / nb, 1) for k, x in zip(names, v)})
end = int(ro[-1] + rl[-1])
print("equal", bool(torch.equal(dst_d[:end], raw_d[:end])))
