"""aircompressor_b200 -- B200-native batched block-compression engine behind the
io.airlift.compress.v3 Compressor/Decompressor API (LZ4, Snappy, Zstandard, XXH64; LZ4 frame format + XXH32, Hadoop block streams).

The product is aircompressor_b200/libaircompress_cuda.so (C ABI in include/aircompress_cuda.h);
this package is the host-side mirror of the reference interface used by tests and bench.py.
"""
from ._native import (F_DEVICE_POINTERS, OP_LZ4_COMPRESS, OP_LZ4_DECOMPRESS, OP_SNAPPY_COMPRESS, OP_SNAPPY_DECOMPRESS,
                      OP_XXH32, OP_XXH64, OP_ZSTD_COMPRESS, OP_ZSTD_DECOMPRESS, NativeLibraryMissing, lib)
from .multi import MultiDeviceEngine
from .api import (BatchEngine, Compressor, Decompressor, IllegalArgumentException, Lz4CudaCompressor, Lz4CudaDecompressor,
                  MalformedInputException, SnappyCudaCompressor, SnappyCudaDecompressor, XxHash64CudaHasher,
                  ZstdCudaCompressor, ZstdCudaDecompressor)
from .lz4_frame import Lz4FrameCudaCompressor, Lz4FrameCudaDecompressor, XxHash32CudaHasher
from .hadoop_streams import (Lz4HadoopCudaInputStream, Lz4HadoopCudaOutputStream, SnappyHadoopCudaInputStream,
                             SnappyHadoopCudaOutputStream)
