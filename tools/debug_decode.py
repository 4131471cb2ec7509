import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aircompressor_b200 as acb
import bench
from oracle.pyoracle import Oracle
codec = sys.argv[1] if len(sys.argv) > 1 else "lz4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
orc = Oracle(); eng = acb.BatchEngine(0); dev = torch.device("cuda", 0)
wl = bench.build_workload(codec, 64, n, orc, orc.max_threads())
def tiled(packed, off, ln):
    reps, stride, stride_al, offs, lens = bench.tile_index(off, ln, n)
    one = torch.zeros(stride_al, dtype=torch.uint8, device=dev); one[:stride] = torch.from_numpy(packed).to(dev)
    return one.repeat(reps), torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev), offs, lens
raw_d, ro_d, rl_d, ro, rl = tiled(wl["raw"], wl["raw_off"], wl["raw_len"])
src_d, so_d, sl_d, so, sl = tiled(wl["comp"], wl["comp_off"], wl["comp_len"])
dst_d = torch.zeros_like(raw_d); ol = torch.zeros(n, dtype=torch.int64, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
op = {"lz4": 1, "snappy": 3}[codec]
for it in range(2):
    dst_d.zero_()
    eng.run_device(op, src_d.data_ptr(), so_d.data_ptr(), sl_d.data_ptr(), dst_d.data_ptr(), ro_d.data_ptr(), rl_d.data_ptr(), ol.data_ptr(), st.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    diff = (dst_d != raw_d)
    nd = int(diff.sum())
    print("iter", it, "status!=0:", int((st != 0).sum()), "len mismatch:", int((ol != rl_d).sum()), "diff bytes:", nd)
    if nd:
        pos = torch.nonzero(diff)[:, 0].cpu().numpy()
        blk = np.searchsorted(ro, pos, side="right") - 1
        ub, cnt = np.unique(blk, return_counts=True)
        print("bad blocks:", len(ub), "distinct idx mod d:", sorted(set((ub % wl["distinct"]).tolist()))[:40])
        for b in ub[:5]:
            p = pos[blk == b]
            print(" block", b, "mod", b % wl["distinct"], "len", rl[b], "first diff at", p[0] - ro[b], "ndiff", len(p), "last", p[-1] - ro[b],
                  "got", dst_d[p[0]:p[0]+8].cpu().numpy(), "want", raw_d[p[0]:p[0]+8].cpu().numpy())
