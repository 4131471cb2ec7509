"""ctypes view of the LZ4 frame API of the image's liblz4 (1.9.4) -- an independent implementation of the frame format used
as an interoperability witness by the frame tests (the reference's bundled liblz4 cannot serve: its LZ4F entry points call a
namespaced XXH32 that the bundle does not export)."""
import ctypes as C

import numpy as np


class FrameInfo(C.Structure):
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int), ("frameType", C.c_int),
                ("contentSize", C.c_ulonglong), ("dictID", C.c_uint), ("blockChecksumFlag", C.c_int)]


class Preferences(C.Structure):
    _fields_ = [("frameInfo", FrameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint), ("favorDecSpeed", C.c_uint),
                ("reserved", C.c_uint * 3)]


class Lz4fNative:
    def __init__(self):
        # by PATH, not by name: the reference's bundled liblz4 (oracle/_ref, loaded by the refnative fixture) carries the same
        # SONAME, and dlopen("liblz4.so.1") would hand back that already loaded library
        import glob
        paths = glob.glob("/lib/x86_64-linux-gnu/liblz4.so.1*") + glob.glob("/usr/lib/x86_64-linux-gnu/liblz4.so.1*") + glob.glob("/usr/lib64/liblz4.so.1*")
        if not paths:
            raise OSError("no system liblz4 in this image")
        L = C.CDLL(sorted(paths)[0])
        sz, vp = C.c_size_t, C.c_void_p
        L.LZ4F_compressFrameBound.restype, L.LZ4F_compressFrameBound.argtypes = sz, [sz, vp]
        L.LZ4F_compressFrame.restype, L.LZ4F_compressFrame.argtypes = sz, [vp, sz, vp, sz, vp]
        L.LZ4F_isError.restype, L.LZ4F_isError.argtypes = C.c_uint, [sz]
        L.LZ4F_createDecompressionContext.restype, L.LZ4F_createDecompressionContext.argtypes = sz, [C.POINTER(vp), C.c_uint]
        L.LZ4F_freeDecompressionContext.restype, L.LZ4F_freeDecompressionContext.argtypes = sz, [vp]
        L.LZ4F_decompress.restype, L.LZ4F_decompress.argtypes = sz, [vp, vp, C.POINTER(sz), vp, C.POINTER(sz), vp]
        self.L = L

    def compress(self, data, block_size_id=7, block_checksum=False, content_checksum=False, content_size=False, level=0):
        p = Preferences()
        p.frameInfo.blockSizeID, p.frameInfo.blockMode = block_size_id, 1      # independent blocks
        p.frameInfo.contentChecksumFlag, p.frameInfo.blockChecksumFlag = int(content_checksum), int(block_checksum)
        p.frameInfo.contentSize = len(data) if content_size else 0
        p.compressionLevel = level
        src = np.frombuffer(bytes(data), dtype=np.uint8)
        cap = self.L.LZ4F_compressFrameBound(len(data), C.byref(p))
        dst = np.zeros(cap, dtype=np.uint8)
        r = self.L.LZ4F_compressFrame(dst.ctypes.data, cap, src.ctypes.data, len(data), C.byref(p))
        assert not self.L.LZ4F_isError(r)
        return dst[:r].tobytes()

    def decompress(self, frame, max_output):
        """Decodes ONE frame (LZ4F stops at its end); returns the bytes or None on error."""
        ctx = C.c_void_p()
        assert not self.L.LZ4F_isError(self.L.LZ4F_createDecompressionContext(C.byref(ctx), 100))
        try:
            src = np.frombuffer(bytes(frame), dtype=np.uint8)
            dst = np.zeros(max(max_output, 1), dtype=np.uint8)
            sp, dp = 0, 0
            while True:
                ss, ds = C.c_size_t(len(frame) - sp), C.c_size_t(max_output - dp)
                r = self.L.LZ4F_decompress(ctx, dst.ctypes.data + dp, C.byref(ds), src.ctypes.data + sp, C.byref(ss), None)
                if self.L.LZ4F_isError(r):
                    return None
                sp += ss.value
                dp += ds.value
                if r == 0:
                    return dst[:dp].tobytes()
                if ss.value == 0 and ds.value == 0:
                    return None
        finally:
            self.L.LZ4F_freeDecompressionContext(ctx)
