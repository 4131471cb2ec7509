#!/usr/bin/env python
"""BASELINE.json configs[4]: block-size sweep 4 KiB - 1 MiB, LZ4 / Snappy / Zstandard, both directions, device-resident batches
of >= 1 GiB of uncompressed data per GPU, with the reference's CPU path (oracle port of the Java codecs and the reference's
bundled native libraries, one call per block, all host threads) timed beside every point on rank 0.

  python tools/sweep.py [--sizes 4,8,...] [--codecs lz4,snappy,zstd] [--gib 1] > profiles/r2_block_size_sweep.jsonl
  python -m torch.distributed.run --nproc-per-node N ... tools/sweep.py     (one rank per GPU, weak scaling, max over ranks)

One JSON line per (codec, op, block size).  Same timing rules as bench.py (3 warm-ups, CUDA events, barrier on both sides);
every decode run is verified against the original bytes on the device.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4,8,16,32,64,128,256,512,1024")
    ap.add_argument("--codecs", default="lz4,snappy,zstd")
    ap.add_argument("--gib", type=float, default=1.0, help="uncompressed GiB per GPU and point")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import aircompressor_b200 as acb
    from oracle.pyoracle import Oracle, RefNative
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    full_aff = os.sched_getaffinity(0)
    acb.lib().acc_bind_host_thread(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    orc = Oracle()
    try:
        ref = RefNative()
    except Exception:
        ref = None
    eng = acb.BatchEngine(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    peak, _ = bench.hbm_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_rate(fn, unc, threads_all):
        """best of N, N/2, N/4 threads over ~cpu_seconds"""
        best = 0.0
        t = threads_all
        for _ in range(3):
            if t < 1:
                break
            fn(t)
            reps, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < args.cpu_seconds / 3:
                fn(t); reps += 1
            best = max(best, unc * reps / (time.perf_counter() - t0) / bench.GiB)
            t //= 2
        return best

    for kib in [int(x) for x in args.sizes.split(",")]:
        n = max(1, int(args.gib * (1 << 20)) // kib)
        for codec in args.codecs.split(","):
            for opname in ("decompress", "compress"):
                os.sched_setaffinity(0, full_aff)
                threads = len(full_aff)
                r = bench.DeviceRun(acb, eng, orc, dev, codec, opname, kib, n, threads)
                tms, kms = r.time(args.steps, 3, barrier)
                t = torch.tensor([tms], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                value = world * args.steps * r.unc_bytes / (float(t[0]) / 1e3) / bench.GiB
                line = {"codec": codec, "op": opname, "block_kib": kib, "blocks_per_gpu": n, "n_gpus": world, "value": value, "unit": "GiB/s",
                        "kernel_ms": float(np.mean(kms)), "frac": r.unc_bytes / (float(np.mean(kms)) / 1e3) / 1e9 / peak,
                        "ratio": (r.comp_bytes / r.unc_bytes) if r.comp_bytes else None, "distinct_blocks": r.wl["distinct"],
                        "frames": "multi-block frames (blocks of one frame are serially dependent)" if codec == "zstd" and kib > 128 else None}
                if rank == 0 and not args.no_cpu:
                    wl = r.wl
                    ns = min(n, max(256, (256 << 10) // kib))        # bounded CPU sample: ~256 MiB of blocks
                    op = bench.CODEC_OPS[(codec, opname)]
                    _, _, _, soff, slen = bench.tile_index(wl["comp_off"] if opname == "decompress" else wl["raw_off"],
                                                           wl["comp_len"] if opname == "decompress" else wl["raw_len"], ns)
                    base = wl["comp"] if opname == "decompress" else wl["raw"]
                    src = np.tile(np.pad(base, (0, ((len(base) + 255) & ~255) - len(base))), (ns + wl["distinct"] - 1) // wl["distinct"])
                    if opname == "decompress":
                        _, _, _, doff, dcap = bench.tile_index(wl["raw_off"], wl["raw_len"], ns)
                        unc_s = int(dcap.sum())
                    else:
                        b = orc.max_compressed_length(codec, int(slen.max()))
                        doff, dcap = np.arange(ns, dtype=np.int64) * b, np.full(ns, b, dtype=np.int64)
                        unc_s = int(slen.sum())
                    dst = np.zeros(int(doff[-1] + dcap[-1]), dtype=np.uint8)
                    line["cpu_port_GiBps"] = cpu_rate(lambda th: orc.batch(op, src, soff, slen, dst, doff, dcap, threads=th), unc_s, threads)
                    if ref is not None and ref.entry_point(op) is not None:
                        line["cpu_native_GiBps"] = cpu_rate(lambda th: orc.native_batch(op, ref, src, soff, slen, dst, doff, dcap, threads=th), unc_s, threads)
                    line["cpu_threads_max"] = threads
                    line["cpu_sample_blocks"] = ns
                if rank == 0:
                    print(json.dumps(line), flush=True)
                del r
                torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
