"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and the reference's bundled native
libraries (oracle/_ref/*.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (aircompressor_b200) never imports this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

OP_LZ4_COMPRESS, OP_LZ4_DECOMPRESS, OP_SNAPPY_COMPRESS, OP_SNAPPY_DECOMPRESS, OP_ZSTD_COMPRESS, OP_ZSTD_DECOMPRESS, OP_XXH64, OP_XXH32 = range(8)


def _u8(buf):
    """bytes / bytearray / numpy uint8 array -> (ctypes pointer, length, keepalive)."""
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8 and buf.flags["C_CONTIGUOUS"]
        return buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size, buf
    if isinstance(buf, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(buf, dtype=np.uint8)
        return arr.ctypes.data_as(C.POINTER(C.c_uint8)), arr.size, arr
    raise TypeError(type(buf))


class Oracle:
    """The restated reference algorithms (kind = "port")."""

    def __init__(self, path=None):
        path = path or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = self.lib = C.CDLL(path)
        p8, i64, pi64 = C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int64)
        for name in ("lz4", "snappy", "zstd"):
            getattr(L, f"orc_{name}_max_compressed_length").restype = i64
            getattr(L, f"orc_{name}_max_compressed_length").argtypes = [i64]
            f = getattr(L, f"orc_{name}_compress")
            f.restype, f.argtypes = i64, [p8, i64, p8, i64]
            f = getattr(L, f"orc_{name}_decompress")
            f.restype, f.argtypes = i64, [p8, i64, p8, i64, pi64]
        L.orc_snappy_uncompressed_length.restype = i64
        L.orc_snappy_uncompressed_length.argtypes = [p8, i64, pi64]
        L.orc_zstd_decompressed_size.restype = i64
        L.orc_zstd_decompressed_size.argtypes = [p8, i64, pi64]
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [p8, i64, C.c_uint64]
        L.orc_xxh64_long.restype = C.c_uint64
        L.orc_xxh64_long.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_xxh32.restype = C.c_uint32
        L.orc_xxh32.argtypes = [p8, i64, C.c_uint32]
        L.orc_batch.restype = i64
        L.orc_batch.argtypes = [C.c_int32, p8, pi64, pi64, p8, pi64, pi64, pi64, i64, C.c_int32]
        try:   # absent from a liboracle.so built before this entry point existed
            L.orc_native_batch.restype = i64
            L.orc_native_batch.argtypes = [C.c_int32, C.c_void_p, p8, pi64, pi64, p8, pi64, pi64, pi64, i64, C.c_int32]
        except AttributeError:
            pass
        L.orc_max_threads.restype = C.c_int32

    def max_compressed_length(self, codec, n):
        return getattr(self.lib, f"orc_{codec}_max_compressed_length")(n)

    def compress(self, codec, data, cap=None):
        src, n, _k = _u8(data)
        cap = self.max_compressed_length(codec, n) if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = getattr(self.lib, f"orc_{codec}_compress")(src, n, out.ctypes.data_as(C.POINTER(C.c_uint8)), cap)
        if r < 0:
            raise OracleError(-r, 0)
        return out[:r].tobytes()

    def decompress_raw(self, codec, data, cap):
        """returns (result, err_offset, output array) without raising."""
        src, n, _k = _u8(data)
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        off = C.c_int64(0)
        r = getattr(self.lib, f"orc_{codec}_decompress")(src, n, out.ctypes.data_as(C.POINTER(C.c_uint8)), cap, C.byref(off))
        return r, off.value, out

    def decompress(self, codec, data, cap):
        r, off, out = self.decompress_raw(codec, data, cap)
        if r < 0:
            raise OracleError(-r, off)
        return out[:r].tobytes()

    def xxh64(self, data, seed=0):
        src, n, _k = _u8(data)
        return self.lib.orc_xxh64(src, n, seed & 0xFFFFFFFFFFFFFFFF)

    def xxh64_long(self, value, seed=0):
        return self.lib.orc_xxh64_long(value & 0xFFFFFFFFFFFFFFFF, seed & 0xFFFFFFFFFFFFFFFF)

    def xxh32(self, data, seed=0):
        src, n, _k = _u8(data)
        return self.lib.orc_xxh32(src, n, seed & 0xFFFFFFFF)

    def max_threads(self):
        return self.lib.orc_max_threads()

    def batch(self, op, src, src_off, src_len, dst, dst_off, dst_cap, threads=1):
        """numpy arrays in, returns (failures, out_len array)."""
        n = len(src_off)
        out_len = np.zeros(n, dtype=np.int64)
        pi64 = C.POINTER(C.c_int64)
        p8 = C.POINTER(C.c_uint8)
        fails = self.lib.orc_batch(
            op, src.ctypes.data_as(p8), src_off.ctypes.data_as(pi64), src_len.ctypes.data_as(pi64),
            dst.ctypes.data_as(p8) if dst is not None else None,
            dst_off.ctypes.data_as(pi64) if dst_off is not None else None,
            dst_cap.ctypes.data_as(pi64) if dst_cap is not None else None,
            out_len.ctypes.data_as(pi64), n, threads)
        return fails, out_len


    def native_batch(self, op, ref, src, src_off, src_len, dst, dst_off, dst_cap, threads=1):
        """The same per-block loop over the reference's bundled native library (ref = RefNative()); ops 0..5 only.
        Returns (failures, out_len array), or None when that library is not available."""
        fn = ref.entry_point(op)
        if fn is None or not hasattr(self.lib, "orc_native_batch"):
            return None
        n = len(src_off)
        out_len = np.zeros(n, dtype=np.int64)
        pi64 = C.POINTER(C.c_int64)
        p8 = C.POINTER(C.c_uint8)
        fails = self.lib.orc_native_batch(
            op, fn, src.ctypes.data_as(p8), src_off.ctypes.data_as(pi64), src_len.ctypes.data_as(pi64),
            dst.ctypes.data_as(p8), dst_off.ctypes.data_as(pi64), dst_cap.ctypes.data_as(pi64),
            out_len.ctypes.data_as(pi64), n, threads)
        return fails, out_len


class OracleError(Exception):
    def __init__(self, status, offset):
        self.status, self.code, self.reason, self.offset = status, status & 0xFF, status >> 8, offset
        super().__init__(f"oracle error code={self.code} reason={self.reason} offset={offset}")


class RefNative:
    """The reference's bundled native libraries (what its *Native* classes bind through FFM):
    liblz4 1.10.0, libsnappy 1.2.1, libzstd 1.5.6, libxxhash 0.8.3 (kind = "reference").
    Falls back to the system liblz4/libzstd/libxxhash of the image when oracle/_ref is absent."""

    def __init__(self):
        ref = os.path.join(_HERE, "_ref")

        def load(name, sysname):
            p = os.path.join(ref, name)
            try:
                # bundled liblz4 has an unresolved LZ4_XXH32_update -> needs lazy binding
                return C.CDLL(p, mode=os.RTLD_LAZY)
            except OSError:
                if sysname is None:
                    return None
                try:
                    return C.CDLL(sysname, mode=os.RTLD_LAZY)
                except OSError:
                    return None

        self.lz4 = load("liblz4.so", "liblz4.so.1")
        self.snappy = load("libsnappy.so", None)
        self.zstd = load("libzstd.so", "libzstd.so.1")
        self.xxhash = load("libxxhash.so", "libxxhash.so.0")
        vp, sz = C.c_void_p, C.c_size_t
        if self.lz4:
            self.lz4.LZ4_compress_fast.restype = C.c_int
            self.lz4.LZ4_compress_fast.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
            self.lz4.LZ4_decompress_safe.restype = C.c_int
            self.lz4.LZ4_decompress_safe.argtypes = [vp, vp, C.c_int, C.c_int]
            self.lz4.LZ4_compressBound.restype = C.c_int
            self.lz4.LZ4_compressBound.argtypes = [C.c_int]
        if self.snappy:
            self.snappy.snappy_compress.restype = C.c_int
            self.snappy.snappy_compress.argtypes = [vp, sz, vp, C.POINTER(sz)]
            self.snappy.snappy_uncompress.restype = C.c_int
            self.snappy.snappy_uncompress.argtypes = [vp, sz, vp, C.POINTER(sz)]
            self.snappy.snappy_max_compressed_length.restype = sz
            self.snappy.snappy_max_compressed_length.argtypes = [sz]
        if self.zstd:
            self.zstd.ZSTD_compress.restype = sz
            self.zstd.ZSTD_compress.argtypes = [vp, sz, vp, sz, C.c_int]
            self.zstd.ZSTD_decompress.restype = sz
            self.zstd.ZSTD_decompress.argtypes = [vp, sz, vp, sz]
            self.zstd.ZSTD_compressBound.restype = sz
            self.zstd.ZSTD_compressBound.argtypes = [sz]
            self.zstd.ZSTD_isError.restype = C.c_uint
            self.zstd.ZSTD_isError.argtypes = [sz]
        if self.xxhash:
            self.xxhash.XXH64.restype = C.c_uint64
            self.xxhash.XXH64.argtypes = [vp, sz, C.c_uint64]
            self.xxhash.XXH32.restype = C.c_uint32
            self.xxhash.XXH32.argtypes = [vp, sz, C.c_uint32]

    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(C.c_void_p)

    def entry_point(self, op):
        """address of the native function behind batch op 0..5 (lz4 c/d, snappy c/d, zstd c/d), or None"""
        table = {0: (self.lz4, "LZ4_compress_fast"), 1: (self.lz4, "LZ4_decompress_safe"), 2: (self.snappy, "snappy_compress"),
                 3: (self.snappy, "snappy_uncompress"), 4: (self.zstd, "ZSTD_compress"), 5: (self.zstd, "ZSTD_decompress")}
        lib, name = table.get(op, (None, None))
        if lib is None:
            return None
        return C.cast(getattr(lib, name), C.c_void_p).value

    def versions(self):
        out = {}
        try:
            self.lz4.LZ4_versionString.restype = C.c_char_p
            out["lz4"] = self.lz4.LZ4_versionString().decode()
        except Exception:
            pass
        try:
            self.zstd.ZSTD_versionString.restype = C.c_char_p
            out["zstd"] = self.zstd.ZSTD_versionString().decode()
        except Exception:
            pass
        return out

    def compress(self, codec, data, level=3):
        src = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        n = src.size
        if codec == "lz4":
            cap = self.lz4.LZ4_compressBound(n)
            out = np.empty(max(cap, 1), dtype=np.uint8)
            r = self.lz4.LZ4_compress_fast(self._ptr(src), self._ptr(out), n, cap, 1)
            if r <= 0:
                raise RuntimeError("LZ4_compress_fast failed")
            return out[:r].tobytes()
        if codec == "snappy":
            cap = self.snappy.snappy_max_compressed_length(n)
            out = np.empty(max(cap, 1), dtype=np.uint8)
            ln = C.c_size_t(cap)
            st = self.snappy.snappy_compress(self._ptr(src), n, self._ptr(out), C.byref(ln))
            if st != 0:
                raise RuntimeError(f"snappy_compress status {st}")
            return out[:ln.value].tobytes()
        if codec == "zstd":
            cap = self.zstd.ZSTD_compressBound(n)
            out = np.empty(max(cap, 1), dtype=np.uint8)
            r = self.zstd.ZSTD_compress(self._ptr(out), cap, self._ptr(src), n, level)
            if self.zstd.ZSTD_isError(r):
                raise RuntimeError("ZSTD_compress failed")
            return out[:r].tobytes()
        raise ValueError(codec)

    def decompress(self, codec, data, cap):
        src = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.empty(max(cap, 1), dtype=np.uint8)
        if codec == "lz4":
            r = self.lz4.LZ4_decompress_safe(self._ptr(src), self._ptr(out), src.size, cap)
            if r < 0:
                raise RuntimeError(f"LZ4_decompress_safe {r}")
            return out[:r].tobytes()
        if codec == "snappy":
            ln = C.c_size_t(cap)
            st = self.snappy.snappy_uncompress(self._ptr(src), src.size, self._ptr(out), C.byref(ln))
            if st != 0:
                raise RuntimeError(f"snappy_uncompress status {st}")
            return out[:ln.value].tobytes()
        if codec == "zstd":
            r = self.zstd.ZSTD_decompress(self._ptr(out), cap, self._ptr(src), src.size)
            if self.zstd.ZSTD_isError(r):
                raise RuntimeError("ZSTD_decompress failed")
            return out[:r].tobytes()
        raise ValueError(codec)

    def xxh64(self, data, seed=0):
        src = np.frombuffer(bytes(data), dtype=np.uint8)
        return self.xxhash.XXH64(self._ptr(src), src.size, seed & 0xFFFFFFFFFFFFFFFF)

    def xxh32(self, data, seed=0):
        src = np.frombuffer(bytes(data), dtype=np.uint8)
        return self.xxhash.XXH32(self._ptr(src), src.size, seed & 0xFFFFFFFF)
