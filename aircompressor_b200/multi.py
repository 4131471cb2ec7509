"""One batch over several GPUs of a host: blocks are independent, so the batch is cut into contiguous,
byte-balanced ranges (sharding.partition_by_bytes), one per device; every device has its own host thread (bound to the GPU's
NUMA node), its own context and its own streams; there is no collective and no peer copy -- the host concatenates the
result arrays.  This is the multi-device shape a Java caller gets from one CudaContext per device and one thread each."""
import threading

import numpy as np

from . import _native as N
from .api import BatchEngine
from .sharding import partition_by_bytes


class MultiDeviceEngine:
    def __init__(self, devices=None, bind_numa=True):
        n = N.lib().acc_device_count()
        self.devices = list(range(n)) if devices is None else list(devices)
        if not self.devices:
            raise RuntimeError("no CUDA device (there is no CPU fallback)")
        self.bind_numa = bind_numa
        self.engines = [None] * len(self.devices)
        self.nodes = [-1] * len(self.devices)
        self._parallel(self._init_one)

    def _parallel(self, fn, *args):
        errs = []

        def run(k):
            try:
                fn(k, *args)
            except BaseException as ex:  # noqa: BLE001 -- re-raised on the calling thread
                errs.append(ex)
        ts = [threading.Thread(target=run, args=(k,)) for k in range(len(self.devices))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    def _init_one(self, k):
        if self.bind_numa:
            self.nodes[k] = N.lib().acc_bind_host_thread(self.devices[k])   # the context's pinned staging is allocated on this thread
        self.engines[k] = BatchEngine(self.devices[k])

    def plan(self, op, src_len, dst_cap):
        """(begin, end) per device: balanced by the bytes that cross PCIe for that block (input + output window)"""
        weight = np.asarray(src_len, dtype=np.int64) + (0 if dst_cap is None else np.asarray(dst_cap, dtype=np.int64))
        return partition_by_bytes(weight, len(self.devices))

    def run_host(self, op, src, src_off, src_len, dst, dst_off, dst_cap):
        """Same contract as BatchEngine.run_host; block i of the batch lands in out_len[i] / status[i] whichever device ran it."""
        n = len(src_off)
        out_len = np.zeros(n, dtype=np.int64)
        status = np.zeros(n, dtype=np.int32)
        so, sl = np.ascontiguousarray(src_off, dtype=np.int64), np.ascontiguousarray(src_len, dtype=np.int64)
        do = None if dst_off is None else np.ascontiguousarray(dst_off, dtype=np.int64)
        dc = None if dst_cap is None else np.ascontiguousarray(dst_cap, dtype=np.int64)
        ranges = self.plan(op, sl, dc)

        def one(k):
            b, e = ranges[k]
            if e <= b:
                return
            if self.bind_numa:
                N.lib().acc_bind_host_thread(self.devices[k])
            ol, st = self.engines[k].run_host(op, src, so[b:e], sl[b:e], dst, None if do is None else do[b:e], None if dc is None else dc[b:e])
            out_len[b:e] = ol
            status[b:e] = st
        self._parallel(one)
        return out_len, status

    def close(self):
        for e in self.engines:
            if e is not None:
                e.close()
