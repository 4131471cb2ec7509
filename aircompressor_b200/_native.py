"""ctypes binding of libaircompress_cuda.so -- the same C ABI a Java FFM record would bind
(include/aircompress_cuda.h; reference mechanism: internal/NativeLoader.java:66-117).

There is no fallback: if the library is missing, or no GPU is usable, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AIRCOMPRESS_CUDA_LIB") or os.path.join(_HERE, "libaircompress_cuda.so")   # the override is for A/B builds of the same ABI

OP_LZ4_COMPRESS, OP_LZ4_DECOMPRESS, OP_SNAPPY_COMPRESS, OP_SNAPPY_DECOMPRESS, OP_ZSTD_COMPRESS, OP_ZSTD_DECOMPRESS, OP_XXH64, OP_XXH32 = range(8)
F_DEVICE_POINTERS = 1

E_MALFORMED, E_DST_TOO_SMALL, E_ARGUMENT, E_CUDA, E_UNSUPPORTED = 1, 2, 3, 4, 5

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Loads (once) and returns the bound library.  Raises NativeLibraryMissing when the CUDA
    extension has not been built -- the product path never degrades to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C aircompressor_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    pi64, pi32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    sig = {
        "acc_device_count": (i32, []),
        "acc_init": (vp, [i32]),
        "acc_init_error": (i32, []),
        "acc_destroy": (None, [vp]),
        "acc_host_alloc": (vp, [i64]),
        "acc_host_free": (None, [vp]),
        "acc_last_error": (i32, [vp, pi64]),
        "acc_code_name": (C.c_char_p, [i32]),
        "acc_reason_text": (C.c_char_p, [i32]),
        "acc_device_numa_node": (i32, [i32]),
        "acc_bind_host_thread": (i32, [i32]),
        "acc_sm_count": (i32, [vp]),
        "acc_kernel_launches": (i64, [vp]),
        "acc_set_tuning": (i32, [vp, i32, i32]),
        "acc_get_stats": (i32, [vp, pi64, i32]),
        "acc_lz4_compress_bound": (i64, [i64]),
        "acc_snappy_compress_bound": (i64, [i64]),
        "acc_zstd_compress_bound": (i64, [i64]),
        "acc_snappy_uncompressed_length": (i64, [vp, i64, pi64]),
        "acc_zstd_frame_content_size": (i64, [vp, i64, pi64]),
        "acc_xxh64": (i64, [vp, vp, i64, i64]),
        "acc_batch": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64]),
        "acc_xxh64_batch": (i32, [vp, vp, vp, vp, vp, i64, i32, i64]),
        "acc_xxh32": (i32, [vp, vp, i64, i32]),
        "acc_xxh32_batch": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i64]),
    }
    for codec in ("lz4", "snappy", "zstd"):
        for d in ("compress", "decompress"):
            sig[f"acc_{codec}_{d}"] = (i64, [vp, vp, i64, vp, i64])
            sig[f"acc_{codec}_{d}_batch"] = (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64])
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here = the .so does not export what the header declares
        f.restype, f.argtypes = res, args
    L._acc_signatures = sig
    _lib = L
    return L


def exported_symbols():
    """Names the header declares; tests check that the built library exports every one."""
    return sorted(lib()._acc_signatures.keys())
