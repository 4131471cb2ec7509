#!/usr/bin/env python
"""bench.py -- headline benchmark: uncompressed GiB/s of the batched block codecs on B200.

Default workload = BASELINE.json configs[1]: LZ4 block decompress, 64 KiB blocks x 65,536 batch per GPU
(4 GiB uncompressed per step per GPU), inputs resident in HBM.  A step = one pass of the hot path over
that batch.  Prints ONE JSON line (see the task contract); `--impl reference` times the reference's CPU
algorithm (oracle port, all host threads) on a bounded sample of the same workload.

The compressed input streams are produced once, untimed, by the restated reference compressor
(oracle/) because the workload is "streams the reference's compressor emits"; the measured path never
touches oracle code, and every decompressed batch is verified against the original bytes on the device.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import benchdata  # noqa: E402

GiB = float(1 << 30)
CODEC_OPS = {
    ("lz4", "decompress"): 1, ("lz4", "compress"): 0, ("snappy", "decompress"): 3, ("snappy", "compress"): 2,
    ("zstd", "decompress"): 5, ("zstd", "compress"): 4, ("xxh64", "hash"): 6, ("xxh32", "hash"): 7,
}


def host_threads():
    """all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def best_thread_count(run_pass, max_threads):
    """The CPU arm should show the reference at its best on this box: with SMT siblings or a CPU quota, fewer threads than
    the affinity mask offers can be faster.  Times one pass (after one warm-up) at N, N/2, N/4, N/8 threads and returns the
    fastest count."""
    best, best_dt = max_threads, None
    t = max_threads
    tried = set()
    while t >= 1 and len(tried) < 4:
        tried.add(t)
        run_pass(t)
        t0 = time.perf_counter()
        run_pass(t)
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = t, dt
        t //= 2
    return best


def time_native_cpu(orc, op, src, soff, slen, dst, doff, dcap, max_threads, unc_bytes, seconds=4.0):
    """The reference's NATIVE CPU path (its bundled liblz4 / libsnappy / libzstd, which its factories prefer when natives are
    enabled: lz4/Lz4Compressor.java:28-43) on the same sample, one call per block, OpenMP over blocks.  Reported beside the
    port of the Java path; None when the library is not available."""
    try:
        from oracle.pyoracle import RefNative
        ref = RefNative()
        if ref.entry_point(op) is None:
            return None
        run = lambda t: orc.native_batch(op, ref, src, soff, slen, dst, doff, dcap, threads=t)
        fails, _ = run(max_threads)
        if fails:
            return None
        threads = best_thread_count(run, max_threads)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            run(threads)
            reps += 1
        dt = time.perf_counter() - t0
        names = {0: "liblz4", 1: "liblz4", 2: "libsnappy", 3: "libsnappy", 4: "libzstd (level 3)", 5: "libzstd"}
        return {"value": unc_bytes * reps / dt / GiB, "unit": "GiB/s", "cores": threads, "kind": "reference",
                "lib": f"{names[op]} bundled by the reference (oracle/_ref), versions {ref.versions()}"}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "error": repr(ex)[:160]}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_WORKLOADS = {}


def build_workload(codec, block_kib, n_blocks, orc, threads):
    """Returns dict with host arrays: distinct blocks, their compressed streams (reference algorithm), tiling."""
    key = (codec, block_kib)
    if key in _WORKLOADS:
        return dict(_WORKLOADS[key], n=n_blocks)
    wl = _build_workload(codec, block_kib, n_blocks, orc, threads)
    _WORKLOADS[key] = wl
    return wl


def _build_workload(codec, block_kib, n_blocks, orc, threads):
    label, pieces = benchdata.load_pieces()
    blocks = benchdata.cut_blocks(pieces, block_kib * 1024)
    raw, raw_off, raw_len = benchdata.pack(blocks)
    if codec in ("xxh64", "xxh32"):
        return {"label": label, "distinct": len(blocks), "raw": raw, "raw_off": raw_off, "raw_len": raw_len,
                "comp": raw, "comp_off": raw_off, "comp_len": raw_len, "n": n_blocks}
    bound = orc.max_compressed_length(codec, int(raw_len.max()))
    caps = np.full(len(blocks), bound, dtype=np.int64)
    coff = (np.arange(len(blocks), dtype=np.int64) * bound)
    cbuf = np.zeros(int(bound * len(blocks)), dtype=np.uint8)
    op = {"lz4": 0, "snappy": 2, "zstd": 4}[codec]
    fails, clen = orc.batch(op, raw, raw_off, raw_len, cbuf, coff, caps, threads=threads)
    assert fails == 0
    streams = [cbuf[coff[i]:coff[i] + clen[i]] for i in range(len(blocks))]
    comp, comp_off, comp_len = benchdata.pack(streams)
    return {"label": label, "distinct": len(blocks), "raw": raw, "raw_off": raw_off, "raw_len": raw_len,
            "comp": comp, "comp_off": comp_off, "comp_len": comp_len, "n": n_blocks}


def tile_index(off, ln, n):
    """Tiles a packed stream of d distinct blocks to n blocks (block i = distinct block i mod d), every
    repetition getting its own copy of the bytes.  Returns (reps, offsets, lengths, total_bytes)."""
    d = len(off)
    stride = int(off[-1] + ln[-1])
    stride_al = (stride + 255) & ~255
    idx = np.arange(n, dtype=np.int64)
    offs = (idx // d) * stride_al + off[idx % d]
    lens = ln[idx % d]
    reps = (n + d - 1) // d
    return reps, stride, stride_al, offs, lens


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port) on this host's cores."""
    if rank != 0:
        return
    from oracle.pyoracle import Oracle
    orc = Oracle()
    threads = host_threads()
    n_sample = min(args.blocks, args.ref_blocks)
    wl = build_workload(args.codec, args.block_kib, n_sample, orc, threads)
    op = CODEC_OPS[(args.codec, args.op)]
    if args.op == "decompress":
        _, _, _, soff, slen = tile_index(wl["comp_off"], wl["comp_len"], n_sample)
        src = np.tile(np.pad(wl["comp"], (0, ((len(wl["comp"]) + 255) & ~255) - len(wl["comp"]))), (n_sample + wl["distinct"] - 1) // wl["distinct"])
        _, _, _, doff, dcap = tile_index(wl["raw_off"], wl["raw_len"], n_sample)
        unc = int(dcap.sum())
        dst = np.zeros(int(doff[-1] + dcap[-1]), dtype=np.uint8)
    else:
        _, _, _, soff, slen = tile_index(wl["raw_off"], wl["raw_len"], n_sample)
        src = np.tile(np.pad(wl["raw"], (0, ((len(wl["raw"]) + 255) & ~255) - len(wl["raw"]))), (n_sample + wl["distinct"] - 1) // wl["distinct"])
        bound = orc.max_compressed_length(args.codec, int(slen.max()))
        dcap = np.full(n_sample, bound, dtype=np.int64)
        doff = np.arange(n_sample, dtype=np.int64) * bound
        unc = int(slen.sum())
        dst = np.zeros(int(bound * n_sample), dtype=np.uint8)
    threads = best_thread_count(lambda t: orc.batch(op, src, soff, slen, dst, doff, dcap, threads=t), threads)
    for _ in range(args.warmup):
        orc.batch(op, src, soff, slen, dst, doff, dcap, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fails, _ = orc.batch(op, src, soff, slen, dst, doff, dcap, threads=threads)
        assert fails == 0
    dt = time.perf_counter() - t0
    value = unc * args.steps / dt / GiB
    sample = f"{n_sample} blocks x {args.block_kib} KiB ({unc / GiB:.2f} GiB uncompressed) per step"
    native = time_native_cpu(orc, op, src, soff, slen, dst, doff, dcap, host_threads(), unc) if args.op != "hash" else None
    line = {
        "impl": "reference", "metric": f"{args.codec}_{args.op}_uncompressed_GiB_per_s", "value": value, "unit": "GiB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": f"synthetic batch: {wl['label']}",
        "config": {"workload": f"{args.codec} block {args.op}, {args.block_kib} KiB blocks x {args.blocks} batch per GPU",
                   "note": "reference CPU algorithm (C restatement of the Java codec, oracle/), one call per block, OpenMP over blocks; thread count = fastest of N, N/2, N/4, N/8"},
        "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": threads, "kind": "port", "sample": sample, "reference_native": native},
        "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args, wl, orc, op, n, threads):
    """cpu_baseline: the oracle port of the reference's Java path (and, beside it, the reference's native path) timed on this
    host's cores on a bounded sample of the same workload, one call per block."""
    ns = min(n, args.ref_blocks)
    _, _, _, soff, slen = tile_index(wl["comp_off"] if args.op == "decompress" else wl["raw_off"],
                                     wl["comp_len"] if args.op == "decompress" else wl["raw_len"], ns)
    base = wl["comp"] if args.op == "decompress" else wl["raw"]
    src = np.tile(np.pad(base, (0, ((len(base) + 255) & ~255) - len(base))), (ns + wl["distinct"] - 1) // wl["distinct"])
    if args.op == "decompress":
        _, _, _, doff, dcap = tile_index(wl["raw_off"], wl["raw_len"], ns)
        unc_s = int(dcap.sum())
    else:
        b = orc.max_compressed_length(args.codec, int(slen.max()))
        doff, dcap = np.arange(ns, dtype=np.int64) * b, np.full(ns, b, dtype=np.int64)
        unc_s = int(slen.sum())
    dst = np.zeros(int(doff[-1] + dcap[-1]), dtype=np.uint8)
    cpu_threads = best_thread_count(lambda t: orc.batch(op, src, soff, slen, dst, doff, dcap, threads=t), threads)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 8.0:
        orc.batch(op, src, soff, slen, dst, doff, dcap, threads=cpu_threads)
        reps += 1
    dt = time.perf_counter() - t0
    cpu_baseline = {"value": unc_s * reps / dt / GiB, "unit": "GiB/s", "cores": cpu_threads, "kind": "port",
                    "sample": f"{ns} blocks x {args.block_kib} KiB x {reps} passes, {cpu_threads} of {threads} host threads (fastest of N, N/2, N/4, N/8), one call per block"}
    cpu_baseline["reference_native"] = time_native_cpu(orc, op, src, soff, slen, dst, doff, dcap, threads, unc_s)
    return cpu_baseline


class DeviceRun:
    """One (codec, op, block size, batch) workload resident in HBM, and the timing of its kernel launches."""

    def __init__(self, acb, eng, orc, dev, codec, opname, block_kib, n, threads):
        import torch
        self.torch, self.eng, self.orc, self.dev = torch, eng, orc, dev
        self.codec, self.opname, self.n = codec, opname, n
        self.op = CODEC_OPS[(codec, opname)]
        wl = self.wl = build_workload(codec, block_kib, n, orc, threads)

        def to_dev_tiled(packed, off, ln):
            reps, stride, stride_al, offs, lens = tile_index(off, ln, n)
            one = torch.zeros(stride_al, dtype=torch.uint8, device=dev)
            one[:stride] = torch.from_numpy(packed).to(dev)
            buf = one.repeat(reps)
            return buf, torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev), offs, lens

        self.raw_d, self.raw_off_d, self.raw_len_d, self.raw_off_h, self.raw_len_h = to_dev_tiled(wl["raw"], wl["raw_off"], wl["raw_len"])
        self.unc_bytes = int(self.raw_len_h.sum())
        if opname == "hash":
            self.src_d, self.src_off_d, self.src_len_d, self.src_off_h, self.src_len_h = self.raw_d, self.raw_off_d, self.raw_len_d, self.raw_off_h, self.raw_len_h
            self.dst_d = torch.zeros(16, dtype=torch.uint8, device=dev)
            self.dst_off_d = torch.zeros(n, dtype=torch.int64, device=dev)
            self.dst_cap_d = torch.zeros(n, dtype=torch.int64, device=dev)
            self.comp_bytes = 0
        elif opname == "decompress":
            self.src_d, self.src_off_d, self.src_len_d, self.src_off_h, self.src_len_h = to_dev_tiled(wl["comp"], wl["comp_off"], wl["comp_len"])
            self.dst_d = torch.zeros_like(self.raw_d)
            self.dst_off_d, self.dst_cap_d = self.raw_off_d, self.raw_len_d
            self.comp_bytes = int(self.src_len_h.sum())
        else:
            self.src_d, self.src_off_d, self.src_len_d, self.src_off_h, self.src_len_h = self.raw_d, self.raw_off_d, self.raw_len_d, self.raw_off_h, self.raw_len_h
            bound = int(getattr(acb.lib(), f"acc_{codec}_compress_bound")(int(self.raw_len_h.max())))
            self.dst_d = torch.zeros(bound * n, dtype=torch.uint8, device=dev)
            self.dst_off_d = torch.arange(n, dtype=torch.int64, device=dev) * bound
            self.dst_cap_d = torch.full((n,), bound, dtype=torch.int64, device=dev)
            self.comp_bytes = None
        self.out_len_d = torch.zeros(n, dtype=torch.int64, device=dev)
        self.status_d = torch.zeros(n, dtype=torch.int32, device=dev)

    def step(self):
        st = self.torch.cuda.current_stream().cuda_stream
        assert st != 0   # handle 0 would mean "the context's own stream" to acc_batch, which torch events cannot see
        self.eng.run_device(self.op, self.src_d.data_ptr(), self.src_off_d.data_ptr(), self.src_len_d.data_ptr(), self.dst_d.data_ptr(),
                            self.dst_off_d.data_ptr(), self.dst_cap_d.data_ptr(), self.out_len_d.data_ptr(), self.status_d.data_ptr(), self.n, st)

    def verify(self):
        """untimed: all blocks OK and (decode) bytes identical to the originals / (hash) values equal to the oracle's"""
        torch, wl = self.torch, self.wl
        assert int((self.status_d != 0).sum()) == 0, "kernel reported errors"
        if self.opname == "hash":
            d = wl["distinct"]
            got = self.out_len_d[:d].cpu().numpy()
            for i in range(0, d, max(1, d // 16)):
                blk = wl["raw"][wl["raw_off"][i]:wl["raw_off"][i] + wl["raw_len"][i]]
                want = self.orc.xxh64(blk.tobytes(), 0) if self.op == 6 else self.orc.xxh32(blk.tobytes(), 0)
                assert int(got[i]) & 0xFFFFFFFFFFFFFFFF == want, "hash mismatch"
        elif self.opname == "decompress":
            assert bool((self.out_len_d == self.raw_len_d).all())
            end = int(self.raw_off_h[-1] + self.raw_len_h[-1])
            assert torch.equal(self.dst_d[:end], self.raw_d[:end]), "decompressed batch differs from the original bytes"
        else:
            self.comp_bytes = int(self.out_len_d.sum())

    def time(self, steps, warmup, barrier):
        """warm-up + verification, then `steps` launches bracketed by barrier + synchronize; returns (total ms, per-launch ms list)"""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        self.verify()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        launches0 = self.eng.kernel_launches
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for a, b in evs:
            a.record(); self.step(); b.record()
        t_end.record()
        barrier()
        self.timed_launches = self.eng.kernel_launches - launches0
        return t_start.elapsed_time(t_end), [a.elapsed_time(b) for a, b in evs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--codec", default="lz4", choices=["lz4", "snappy", "zstd", "xxh64", "xxh32"])
    ap.add_argument("--op", default="decompress", choices=["compress", "decompress"])
    ap.add_argument("--block-kib", type=int, default=0, help="default 64 (lz4/snappy) or 128 (zstd)")
    ap.add_argument("--blocks", type=int, default=0, help="default: 4 GiB of uncompressed data per GPU")
    ap.add_argument("--ref-blocks", type=int, default=8192, help="bounded sample for the CPU arms")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-extra", action="store_true", help="skip the short per-codec side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--decode-path", type=int, default=0, help="LZ4 / Snappy decode: 0 library default, 1 step decoder, 2 record path")
    ap.add_argument("--pipeline", type=int, default=0, help="host-pointer path: 0 auto, 1 single pass, k>1 overlapped runs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: ONE batch of --blocks blocks is split over the ranks (contiguous, byte-balanced: BASELINE.md s3 config 3 wording)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--profile", action="store_true", help="profiling run (under ncu): no e2e, no cpu baseline, warm-up as given")
    args = ap.parse_args()
    if args.codec in ("xxh64", "xxh32"):
        args.op = "hash"
    if args.block_kib == 0:
        args.block_kib = 128 if args.codec == "zstd" else 64
    if args.blocks == 0:
        args.blocks = (4 << 20) // args.block_kib
    if args.impl == "cuda" and not args.profile:
        args.warmup = max(args.warmup, 3)
    if args.profile:
        args.no_cpu_baseline = True

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import aircompressor_b200 as acb
    from oracle.pyoracle import Oracle  # input preparation + cpu_baseline leg only

    # host side of this rank next to its GPU: CPU affinity first, so that everything pinned below is placed on that node
    full_affinity = os.sched_getaffinity(0)
    numa_node = -1 if args.no_numa_bind else int(acb.lib().acc_bind_host_thread(local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    orc = Oracle()
    threads = host_threads()
    eng = acb.BatchEngine(local_rank)
    if args.ctas_per_sm:
        eng.set_tuning(0, args.ctas_per_sm)
    if args.decode_path:
        eng.set_tuning(1, args.decode_path)
    op = CODEC_OPS[(args.codec, args.op)]
    n = args.blocks
    if args.scaling == "strong" and world > 1:
        # one batch, contiguous split: the tiling makes every block of a residue class the same size, so an even split of the
        # block range is the byte-balanced split of sharding.partition_by_bytes up to one tile
        n = args.blocks // world + (1 if rank < args.blocks % world else 0)

    # ---------------- workload (weak scaling: every rank owns a full batch) ----------------
    # all device work of the benchmark runs on one explicit stream; its handle is what the C ABI gets
    bench_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(bench_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    run = DeviceRun(acb, eng, orc, dev, args.codec, args.op, args.block_kib, n, threads)
    wl = run.wl
    unc_bytes = run.unc_bytes
    src_d, dst_d, out_len_d = run.src_d, run.dst_d, run.out_len_d
    src_off_h, src_len_h, dst_off_d, dst_cap_d = run.src_off_h, run.src_len_h, run.dst_off_d, run.dst_cap_d

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, kernel_ms = run.time(args.steps, args.warmup, barrier)
    clocks = sampler.stop() if rank == 0 else None
    launches = run.timed_launches
    comp_bytes = run.comp_bytes
    total_ms = max_over_ranks(total_ms)
    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    job_unc_bytes = sum_over_ranks(unc_bytes)        # weak: world x the rank's batch; strong: the one batch
    value = args.steps * job_unc_bytes / (total_ms / 1e3) / GiB

    # ---------------- e2e: host buffers through the C ABI (H2D + kernel + D2H inside the timed region) ----------------
    e2e = None
    try:
        if args.profile or args.op == "hash" or args.e2e_steps <= 0:
            raise RuntimeError("skipped")
        h_src = torch.empty(src_d.numel(), dtype=torch.uint8, pin_memory=True)
        h_src.copy_(src_d)
        h_dst = torch.empty(dst_d.numel(), dtype=torch.uint8, pin_memory=True)
        hs, hd = h_src.numpy(), h_dst.numpy()
        so_h, sl_h = src_off_h, src_len_h
        do_h, dc_h = dst_off_d.cpu().numpy(), dst_cap_d.cpu().numpy()
        # what the PCIe link of this GPU gives in each direction alone (pinned memory, one cudaMemcpyAsync), and the e2e value
        # that bound allows when upload and download overlap perfectly: bytes / max(t_h2d, t_d2h)
        def copy_gbps(dst_t, src_t, reps=3):
            best = None
            for _ in range(reps):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); dst_t.copy_(src_t, non_blocking=True); b_.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b_)
                best = ms if best is None or ms < best else best
            return src_t.numel() / (best / 1e3) / 1e9, best
        h2d_gbps, h2d_ms = copy_gbps(src_d, h_src)
        d2h_gbps, d2h_ms = copy_gbps(h_dst, dst_d)
        # both directions at once (what the overlapped path asks of the link): upload on one stream, download on another
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        with torch.cuda.stream(s_up):
            ev[0].record(); src_d.copy_(h_src, non_blocking=True); ev[1].record()
        with torch.cuda.stream(s_dn):
            ev[2].record(); h_dst.copy_(dst_d, non_blocking=True); ev[3].record()
        torch.cuda.synchronize()
        dup_up_ms, dup_dn_ms = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
        link = {"h2d_GBps": h2d_gbps, "d2h_GBps": d2h_gbps, "numa_node": numa_node,
                "overlap_bound_GiBps": unc_bytes / (max(h2d_ms, d2h_ms) / 1e3) / GiB,
                "duplex_h2d_GBps": h_src.numel() / (dup_up_ms / 1e3) / 1e9, "duplex_d2h_GBps": dst_d.numel() / (dup_dn_ms / 1e3) / 1e9,
                "duplex_bound_GiBps": unc_bytes / (max(dup_up_ms, dup_dn_ms) / 1e3) / GiB}

        def timed_host_calls(pipeline):
            eng.set_tuning(3, pipeline)
            eng.run_host(op, hs, so_h, sl_h, hd, do_h, dc_h)  # warm staging allocations
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                olen, stt = eng.run_host(op, hs, so_h, sl_h, hd, do_h, dc_h)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            assert (stt == 0).all()
            return float(tt[0]), olen

        dt_single, _ = timed_host_calls(1) if args.pipeline == 0 else (None, None)
        h_dst.zero_()
        dt, olen = timed_host_calls(args.pipeline)
        eng.set_tuning(3, 0)
        # the host-path result must equal what the device-resident path produced
        dev_len = out_len_d.cpu().numpy()
        assert np.array_equal(olen, dev_len), "e2e lengths differ from the device-resident run"   # encoders are deterministic
        if args.op == "decompress":
            end = int(do_h[-1] + dc_h[-1])
            assert torch.equal(h_dst[:end], dst_d[:end].cpu()), "e2e output differs from the device-resident run"
        h2d = int(src_off_h[-1] + src_len_h[-1]) + 4 * 8 * n
        d2h = int(do_h[-1] + dc_h[-1]) + 12 * n
        e2e = {"value": args.e2e_steps * job_unc_bytes / dt / GiB, "unit": "GiB/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "path": "acc_batch with pinned host buffers; upload, kernels and download of consecutive runs of blocks overlap on three streams, one sync per call",
               "single_pass_value": (args.e2e_steps * job_unc_bytes / dt_single / GiB) if dt_single else None,
               "link": link}
        del h_src, h_dst
    except Exception as ex:  # noqa: BLE001
        e2e = {"value": None, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:200]}

    # ---------------- single-block calls: the Java-shaped entry points (acc_<codec>_<op>: pageable host memory, one block) ----------------
    single = None
    if rank == 0 and not args.profile and args.op != "hash":
        try:
            L = acb.lib()
            i0 = wl["distinct"] // 2
            blk = np.ascontiguousarray(wl["raw"][wl["raw_off"][i0]:wl["raw_off"][i0] + wl["raw_len"][i0]])
            stream = np.ascontiguousarray(wl["comp"][wl["comp_off"][i0]:wl["comp_off"][i0] + wl["comp_len"][i0]])
            cbound = int(getattr(L, f"acc_{args.codec}_compress_bound")(blk.size))
            cbuf, dbuf = np.zeros(cbound, dtype=np.uint8), np.zeros(blk.size, dtype=np.uint8)
            fc, fd = getattr(L, f"acc_{args.codec}_compress"), getattr(L, f"acc_{args.codec}_decompress")
            h = eng._ctx.handle

            def lat(f, a, alen, b, blen, reps=200):
                ts = []
                for _ in range(reps + 20):
                    t0 = time.perf_counter()
                    r = f(h, a.ctypes.data, alen, b.ctypes.data, blen)
                    ts.append(time.perf_counter() - t0)
                    assert r > 0
                return float(np.median(ts[20:]) * 1e6)
            d_us = lat(fd, stream, stream.size, dbuf, blk.size)
            assert np.array_equal(dbuf, blk)
            c_us = lat(fc, blk, blk.size, cbuf, cbound)
            single = {"block_bytes": int(blk.size), "decompress_us": d_us, "compress_us": c_us, "calls": 200,
                      "path": f"acc_{args.codec}_decompress / _compress: pinned staging, upload, kernel, download, one stream sync per call (median)"}
        except Exception as ex:  # noqa: BLE001
            single = {"error": repr(ex)[:160]}

    # ---------------- extras: the other codec-directions of BASELINE.json configs[2..3] at this N ----------------
    # 1 GiB of uncompressed data per GPU each (Snappy / LZ4 at 64 KiB, Zstandard at 128 KiB), device-resident, same timing
    # rules as the headline (3 warm-ups, CUDA events, barrier on both sides, max over ranks); every run verifies itself.
    peak, peak_src = hbm_peak()
    extras = None
    if not args.no_extra and not args.profile and args.op != "hash":
        del run, src_d, dst_d, out_len_d, dst_off_d, dst_cap_d
        torch.cuda.empty_cache()
        extras = {}
        for codec, opname, kib in (("lz4", "decompress", 64), ("lz4", "compress", 64), ("snappy", "decompress", 64), ("snappy", "compress", 64),
                                   ("zstd", "decompress", 128), ("zstd", "compress", 128)):
            if (codec, opname, kib) == (args.codec, args.op, args.block_kib):
                continue
            try:
                nb = (1 << 20) // kib
                r = DeviceRun(acb, eng, orc, dev, codec, opname, kib, nb, threads)
                x_steps = 5
                tms, kms = r.time(x_steps, 3, barrier)
                tms = max_over_ranks(tms)
                extras[f"{codec}_{opname}"] = {
                    "value": world * x_steps * r.unc_bytes / (tms / 1e3) / GiB, "unit": "GiB/s", "block_kib": kib, "blocks_per_gpu": nb,
                    "steps": x_steps, "warmup": 3, "ratio": (r.comp_bytes / r.unc_bytes) if r.comp_bytes else None,
                    "roofline_frac": r.unc_bytes / (float(np.mean(kms)) / 1e3) / 1e9 / peak,
                    "frac_min_traffic": (r.unc_bytes + (r.comp_bytes or 0)) / (float(np.mean(kms)) / 1e3) / 1e9 / peak}
                del r
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001 -- an extra must never cost the headline line
                extras[f"{codec}_{opname}"] = {"value": None, "error": repr(ex)[:160]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    avg_kernel_ms = float(np.mean(kernel_ms))
    achieved = unc_bytes / (avg_kernel_ms / 1e3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            t = json.load(open(prof)).get(f"{args.codec}_{args.op}")
            # ncu capture of this kernel on the same workload shape; scaled to this run's batch size
            traffic = t["dram_bytes_per_launch"] * n / t["captured_with_blocks"] if t else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": unc_bytes,
                "min_traffic_bytes_per_launch": unc_bytes + (comp_bytes or 0),
                "frac_min_traffic": (unc_bytes + (comp_bytes or 0)) / (avg_kernel_ms / 1e3) / 1e9 / peak,
                "kernel_ms_avg": avg_kernel_ms, "kernel_ms_min": float(np.min(kernel_ms))}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and args.op != "hash":
        os.sched_setaffinity(0, full_affinity)     # the CPU arm gets every core of the host again
        threads = host_threads()
        try:
            cpu_baseline = cpu_baseline_leg(args, wl, orc, op, n, threads)
        except Exception as ex:  # noqa: BLE001 -- never lose the GPU line because the CPU leg failed
            cpu_baseline = {"value": None, "unit": "GiB/s", "cores": 0, "kind": "port", "sample": "failed", "error": repr(ex)[:200]}

    line = {
        "metric": f"{args.codec}_{args.op}_uncompressed_GiB_per_s", "value": value, "unit": "GiB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": f"synthetic batch: {wl['label']}, tiled",
        "config": {"workload": f"{args.codec} block {args.op}, {args.block_kib} KiB blocks x {n} batch per GPU" + (" (BASELINE.json configs[1])" if (args.codec, args.op, args.block_kib, n) == ("lz4", "decompress", 64, 65536) else ""),
                   "distinct_blocks": wl["distinct"], "uncompressed_bytes_per_gpu": unc_bytes, "compressed_bytes_per_gpu": comp_bytes,
                   "ratio": (comp_bytes / unc_bytes) if comp_bytes else None,
                   "parallelism": (f"independent blocks, batch per GPU x{world}, no collective" if args.scaling == "weak" else f"one batch of {args.blocks} blocks split contiguously over {world} GPUs, no collective"),
                   "l2": "inputs+outputs (>= 4 GiB per step) far exceed the 126 MB L2; no flush needed",
                   "input_streams": "reference algorithm (oracle port of Lz4RawCompressor etc.), prepared untimed"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "single_block": single, "extras": extras, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
