"""MultiDeviceEngine (aircompressor_b200/multi.py): one batch cut into contiguous byte-balanced ranges, one host thread + context
per device.  Runs on however many GPUs the box has (1 is the degenerate case; `gpurun --gpus 2` exercises the split): results
must equal the single-device run block for block, and the split must cover the batch exactly."""
import numpy as np
import pytest

import aircompressor_b200 as acb
import benchdata

pytestmark = pytest.mark.gpu


def test_multi_device_equals_single_device(engine, oracle, pieces):
    blocks = benchdata.cut_blocks(pieces, 64 * 1024)[:600]
    src, so, sl = benchdata.pack(blocks)
    multi = acb.MultiDeviceEngine()
    try:
        assert len(multi.devices) == acb.lib().acc_device_count() >= 1
        for codec, cop, dop in (("lz4", acb.OP_LZ4_COMPRESS, acb.OP_LZ4_DECOMPRESS), ("snappy", acb.OP_SNAPPY_COMPRESS, acb.OP_SNAPPY_DECOMPRESS),
                                ("zstd", acb.OP_ZSTD_COMPRESS, acb.OP_ZSTD_DECOMPRESS)):
            bound = getattr(acb.lib(), f"acc_{codec}_compress_bound")(int(sl.max()))
            caps = np.full(len(blocks), bound, dtype=np.int64)
            do = np.arange(len(blocks), dtype=np.int64) * bound
            ranges = multi.plan(cop, sl, caps)
            assert ranges[0][0] == 0 and ranges[-1][1] == len(blocks) and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            comp_m, comp_s = np.zeros(int(bound * len(blocks)), dtype=np.uint8), np.zeros(int(bound * len(blocks)), dtype=np.uint8)
            clen_m, st_m = multi.run_host(cop, src, so, sl, comp_m, do, caps)
            clen_s, st_s = engine.run_host(cop, src, so, sl, comp_s, do, caps)
            assert (st_m == 0).all() and np.array_equal(clen_m, clen_s)
            for i in range(0, len(blocks), 37):
                assert np.array_equal(comp_m[do[i]:do[i] + clen_m[i]], comp_s[do[i]:do[i] + clen_s[i]]), (codec, i)   # encoders are deterministic
            back = np.zeros_like(src)
            dlen, st = multi.run_host(dop, comp_m, do, clen_m, back, so, sl)
            assert (st == 0).all() and np.array_equal(dlen, sl) and np.array_equal(back, src), codec
        h_m, _ = multi.run_host(acb.OP_XXH64, src, so, sl, None, None, None)
        assert all(int(h_m[i]) & 0xFFFFFFFFFFFFFFFF == oracle.xxh64(blocks[i].tobytes(), 0) for i in range(0, len(blocks), 23))
    finally:
        multi.close()


def test_numa_binding_is_harmless():
    import threading
    res = {}

    def worker():      # the binding is per thread: keep it off the test runner's main thread
        res["node"] = acb.lib().acc_device_numa_node(0)
        res["bound"] = acb.lib().acc_bind_host_thread(0)
    t = threading.Thread(target=worker)
    t.start(); t.join()
    assert res["bound"] in (-1, res["node"])
