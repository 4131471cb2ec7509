"""LZ4 frame codec (row f1) end to end: one user call = one frame of 4 MiB blocks = one GPU batch.  Wall clock over host buffers
(framing in Python/numpy, blocks and checksums through the C ABI), round trip verified; prints one JSON line.
usage: python tools/frame_bench.py [MiB]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aircompressor_b200 as acb  # noqa: E402
import benchdata  # noqa: E402


def main(mib):
    _label, pieces = benchdata.load_pieces()
    corpus = np.concatenate(pieces)
    data = np.tile(corpus, (mib << 20) // len(corpus) + 1)[:mib << 20].copy()
    comp, dec = acb.Lz4FrameCudaCompressor(), acb.Lz4FrameCudaDecompressor()
    frame = np.zeros(comp.maxCompressedLength(len(data)), dtype=np.uint8)
    back = np.zeros(len(data), dtype=np.uint8)
    res = {}
    for rep in range(3):
        t0 = time.perf_counter(); n = comp.compress(data, 0, len(data), frame, 0, len(frame)); t1 = time.perf_counter()
        m = dec.decompress(frame, 0, n, back, 0, len(back)); t2 = time.perf_counter()
        assert m == len(data) and np.array_equal(back, data)
        res = {"compress_gib_s": len(data) / (t1 - t0) / 2**30, "decompress_gib_s": len(data) / (t2 - t1) / 2**30}
    print(json.dumps({"metric": "lz4 frame codec, host buffers, wall clock", "input_mib": mib, "blocks": (len(data) + (4 << 20) - 1) // (4 << 20),
                      "ratio": n / len(data), **{k: round(v, 3) for k, v in res.items()}}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
