"""Host-side mirror of the reference's block codec API (io.airlift.compress.v3.Compressor /
Decompressor, M/Compressor.java:18-36, M/Decompressor.java:18-31) over the C ABI.

Class and method names, argument meaning and error behaviour follow the reference so the parity
tests read like AbstractTestCompression.java.  The JVM is not available in this image, so this
Python layer plays the role of the Java shim classes shown in INTEGRATION.md
(Lz4CudaCompressor, Lz4CudaDecompressor, SnappyCuda*, ZstdCuda*, XxHash64CudaHasher): argument
validation happens here exactly as in the Java wrappers (Lz4JavaCompressor.verifyRange,
lz4/Lz4JavaCompressor.java:78-84), everything else is one FFM-shaped downcall.
"""
import ctypes as C

import numpy as np

from . import _native as N


class MalformedInputException(RuntimeError):
    """io.airlift.compress.v3.MalformedInputException (MalformedInputException.java:21-35)."""

    def __init__(self, offset, reason="Malformed input"):
        super().__init__(f"{reason}: offset={offset}")
        self.offset = offset
        self.reason = reason


class IllegalArgumentException(ValueError):
    pass


def _as_array(buf, writable=False):
    if buf is None:
        raise TypeError("data is null")  # requireNonNull -> NullPointerException in Java
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or not buf.flags["C_CONTIGUOUS"]:
            raise TypeError("expected a contiguous uint8 array")
        return buf
    arr = np.frombuffer(buf, dtype=np.uint8)
    if writable and not arr.flags["WRITEABLE"]:
        raise TypeError("output buffer is read-only")
    return arr


def _verify_range(arr, offset, length):
    # Lz4JavaCompressor.verifyRange (lz4/Lz4JavaCompressor.java:78-84)
    if offset < 0 or length < 0 or offset + length > arr.size:
        raise IllegalArgumentException(f"Invalid offset or length ({offset}, {length}) in array of length {arr.size}")


def _raise_for(status, offset, dst_cap=None, bound=None):
    code, reason = status & 0xFF, status >> 8
    text = N.lib().acc_reason_text(reason).decode()
    if code == N.E_MALFORMED:
        raise MalformedInputException(offset, text)
    if code == N.E_DST_TOO_SMALL:
        raise IllegalArgumentException(f"Output buffer too small: {text}")
    if code == N.E_ARGUMENT:
        raise IllegalArgumentException(text)
    if code == N.E_UNSUPPORTED:
        raise NotImplementedError(f"unsupported by this build: {text}")
    raise RuntimeError(f"{N.lib().acc_code_name(code).decode()}: reason={reason}")


class _Context:
    """One acc_ctx per codec object, like the scratch each Java codec instance owns."""

    def __init__(self, device=0):
        L = N.lib()
        self._L = L
        self.handle = L.acc_init(device)
        if not self.handle:
            st = L.acc_init_error()
            raise RuntimeError(f"acc_init({device}) failed: {L.acc_code_name(st & 0xFF).decode()} (cudaError {st >> 8}); "
                               "no CPU fallback exists")
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self._L.acc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Codec:
    codec = None

    def __init__(self, device=0):
        self._ctx = _Context(device)
        self._L = N.lib()

    def close(self):
        self._ctx.close()

    def _call(self, direction, inp, in_off, in_len, out, out_off, out_cap):
        src = _as_array(inp)
        dst = _as_array(out, writable=True)
        _verify_range(src, in_off, in_len)
        _verify_range(dst, out_off, out_cap)
        f = getattr(self._L, f"acc_{self.codec}_{direction}")
        r = f(self._ctx.handle, src.ctypes.data + in_off, in_len, dst.ctypes.data + out_off, out_cap)
        if r < 0:
            off = C.c_int64(0)
            st = self._L.acc_last_error(self._ctx.handle, C.byref(off))
            self._on_error(st, off.value)
        return int(r)

    def _on_error(self, status, offset):
        _raise_for(status, offset)


class Compressor(_Codec):
    """Compressor.java:18-36."""

    def maxCompressedLength(self, uncompressed_size):
        return int(getattr(self._L, f"acc_{self.codec}_compress_bound")(uncompressed_size))

    def compress(self, input, inputOffset=None, inputLength=None, output=None, outputOffset=None, maxOutputLength=None):
        if output is None and inputLength is None:  # compress(MemorySegment input, MemorySegment output)
            output, inputOffset = inputOffset, 0
        if inputLength is None:
            src, dst = _as_array(input), _as_array(output, writable=True)
            return self._call("compress", src, 0, src.size, dst, 0, dst.size)
        return self._call("compress", input, inputOffset, inputLength, output, outputOffset, maxOutputLength)

    def getRetainedSizeInBytes(self, input_length=0):
        return 0


class Decompressor(_Codec):
    """Decompressor.java:18-31."""

    def decompress(self, input, inputOffset=None, inputLength=None, output=None, outputOffset=None, maxOutputLength=None):
        if output is None and inputLength is None:  # decompress(MemorySegment input, MemorySegment output)
            output, inputOffset = inputOffset, 0
        if inputLength is None:
            src, dst = _as_array(input), _as_array(output, writable=True)
            return self._call("decompress", src, 0, src.size, dst, 0, dst.size)
        return self._call("decompress", input, inputOffset, inputLength, output, outputOffset, maxOutputLength)


class Lz4CudaCompressor(Compressor):
    codec = "lz4"


class Lz4CudaDecompressor(Decompressor):
    codec = "lz4"

    def _on_error(self, status, offset):
        # Lz4RawDecompressor.java:52-57: zero-capacity output with a non-{0} stream returns -1, no exception
        if (status >> 8) == 6:
            raise _Lz4MinusOne()
        _raise_for(status, offset)

    def _call(self, *a):
        try:
            return super()._call(*a)
        except _Lz4MinusOne:
            return -1


class _Lz4MinusOne(Exception):
    pass


class SnappyCudaCompressor(Compressor):
    codec = "snappy"


class SnappyCudaDecompressor(Decompressor):
    codec = "snappy"

    def getUncompressedLength(self, compressed, compressedOffset=0):
        """SnappyDecompressor.getUncompressedLength (snappy/SnappyDecompressor.java:22)."""
        src = _as_array(compressed)
        _verify_range(src, compressedOffset, 0)      # the reference bounds the varint read by compressed.length
        off = C.c_int64(0)
        r = self._L.acc_snappy_uncompressed_length(src.ctypes.data + compressedOffset, src.size - compressedOffset, C.byref(off))
        if r < 0:
            _raise_for(int(-r), off.value)
        return int(r)


class ZstdCudaCompressor(Compressor):
    codec = "zstd"


class ZstdCudaDecompressor(Decompressor):
    codec = "zstd"

    def getDecompressedSize(self, input, offset, length):
        """ZstdDecompressor.getDecompressedSize (zstd/ZstdDecompressor.java:21)."""
        src = _as_array(input)
        _verify_range(src, offset, length)
        off = C.c_int64(0)
        r = self._L.acc_zstd_frame_content_size(src.ctypes.data + offset, length, C.byref(off))
        if r == -1:
            return -1  # frame does not record its content size (FrameHeader.contentSize == -1 in the Java)
        if r < 0:
            _raise_for(int(-r), off.value)
        return int(r)


class XxHash64CudaHasher:
    """One-shot XXH64 (XxHash64Hasher.hash overloads, xxhash/XxHash64Hasher.java:44-80)."""

    def __init__(self, device=0):
        self._ctx = _Context(device)
        self._L = N.lib()

    def hash(self, input, offset=0, length=None, seed=0):
        src = _as_array(input)
        if length is None:
            length = src.size - offset
        _verify_range(src, offset, length)
        h = self._L.acc_xxh64(self._ctx.handle, src.ctypes.data + offset, length, C.c_int64(_to_signed(seed)).value)
        return h & 0xFFFFFFFFFFFFFFFF

    def hashLong(self, value, seed=0):
        """hash(long value[, long seed]) (XxHash64Hasher.java:44-53): the 8 bytes of `value` in little-endian order."""
        return self.hash((value & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "little"), 0, 8, seed)

    def close(self):
        self._ctx.close()


def _to_signed(v):
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


class BatchEngine:
    """Batched entry points: n independent blocks per call (acc_batch).  Works on host numpy arrays
    (the library stages host<->device itself) or on device buffers given as raw pointers
    (torch tensors' data_ptr()) with ACC_F_DEVICE_POINTERS, asynchronously on `stream`."""

    def __init__(self, device=0):
        self._ctx = _Context(device)
        self._L = N.lib()
        self.device = device

    @property
    def sm_count(self):
        return self._L.acc_sm_count(self._ctx.handle)

    @property
    def kernel_launches(self):
        return int(self._L.acc_kernel_launches(self._ctx.handle))

    def set_tuning(self, key, value):
        return self._L.acc_set_tuning(self._ctx.handle, key, value)

    STAT_NAMES = ("batches", "blocks", "launches", "host_calls", "h2d_bytes", "d2h_bytes", "last_call_us", "total_call_us")

    def stats(self):
        """acc_get_stats: counters of this context since it was created (include/aircompress_cuda.h ACC_STAT_*)."""
        buf = (C.c_int64 * len(self.STAT_NAMES))()
        n = self._L.acc_get_stats(self._ctx.handle, buf, len(self.STAT_NAMES))
        return {self.STAT_NAMES[i]: int(buf[i]) for i in range(n)}

    def run_host(self, op, src, src_off, src_len, dst, dst_off, dst_cap):
        """numpy in/out. Returns (out_len, status) arrays."""
        n = len(src_off)
        out_len = np.zeros(n, dtype=np.int64)
        status = np.zeros(n, dtype=np.int32)
        so, sl = np.ascontiguousarray(src_off, dtype=np.int64), np.ascontiguousarray(src_len, dtype=np.int64)
        if op in (N.OP_XXH64, N.OP_XXH32):
            r = self._L.acc_batch(self._ctx.handle, op, src.ctypes.data, so.ctypes.data, sl.ctypes.data, None, None, None,
                                  out_len.ctypes.data, status.ctypes.data, n, 0, 0)
        else:
            do, dc = np.ascontiguousarray(dst_off, dtype=np.int64), np.ascontiguousarray(dst_cap, dtype=np.int64)
            r = self._L.acc_batch(self._ctx.handle, op, src.ctypes.data, so.ctypes.data, sl.ctypes.data, dst.ctypes.data,
                                  do.ctypes.data, dc.ctypes.data, out_len.ctypes.data, status.ctypes.data, n, 0, 0)
        if r != 0:
            raise RuntimeError(f"acc_batch failed: status {-r:#x}")
        return out_len, status

    def run_device(self, op, src_ptr, src_off_ptr, src_len_ptr, dst_ptr, dst_off_ptr, dst_cap_ptr, out_len_ptr, status_ptr, n, stream=0):
        r = self._L.acc_batch(self._ctx.handle, op, src_ptr, src_off_ptr, src_len_ptr, dst_ptr, dst_off_ptr, dst_cap_ptr,
                              out_len_ptr, status_ptr, n, N.F_DEVICE_POINTERS, stream)
        if r != 0:
            raise RuntimeError(f"acc_batch failed: status {-r:#x}")

    def close(self):
        self._ctx.close()
