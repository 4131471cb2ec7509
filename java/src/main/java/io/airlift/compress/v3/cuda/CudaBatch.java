/*
 * Licensed under the Apache License, Version 2.0 (the "License");
 * you may not use this file except in compliance with the License.
 * You may obtain a copy of the License at
 *
 *     http://www.apache.org/licenses/LICENSE-2.0
 *
 * Unless required by applicable law or agreed to in writing, software
 * distributed under the License is distributed on an "AS IS" BASIS,
 * WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND, either express or implied.
 * See the License for the specific language governing permissions and
 * limitations under the License.
 */
package io.airlift.compress.v3.cuda;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

/**
 * The GPU-shaped entry point: n independent blocks in one call (acc_batch).  A caller that has many chunks -- the Hadoop
 * block streams submit one 256 KiB chunk at a time today (lz4/Lz4HadoopOutputStream.java:107-117) -- collects them in one
 * source segment and submits them together; block i reads src[srcOffset[i], srcOffset[i] + srcLength[i]) and writes at most
 * dstCapacity[i] bytes at dst + dstOffset[i].
 *
 * Source and destination should be native segments from {@link #allocatePinned(long)} (copies then run at PCIe speed and
 * large batches overlap upload, kernels and download); pageable native segments work but serialise the copies.
 * Output contract: bytes [0, outLength[i]) of window i are the result; the rest of a window is unspecified; nothing outside
 * the windows is written.  A failed block has status[i] != 0 (same status words as the single-block classes; outLength[i]
 * then holds the error offset) and does not affect its neighbours.
 */
public final class CudaBatch
        implements AutoCloseable
{
    public static final int LZ4_COMPRESS = AircompressCuda.OP_LZ4_COMPRESS;
    public static final int LZ4_DECOMPRESS = AircompressCuda.OP_LZ4_DECOMPRESS;
    public static final int SNAPPY_COMPRESS = AircompressCuda.OP_SNAPPY_COMPRESS;
    public static final int SNAPPY_DECOMPRESS = AircompressCuda.OP_SNAPPY_DECOMPRESS;
    public static final int ZSTD_COMPRESS = AircompressCuda.OP_ZSTD_COMPRESS;
    public static final int ZSTD_DECOMPRESS = AircompressCuda.OP_ZSTD_DECOMPRESS;
    public static final int XXH64 = AircompressCuda.OP_XXH64;

    public record Result(long[] outLength, int[] status) {}

    private final CudaContext context;

    public CudaBatch()
    {
        this(0);
    }

    public CudaBatch(int device)
    {
        this.context = new CudaContext(device);
    }

    /** pinned host memory; release with {@link #freePinned(MemorySegment)} */
    public static MemorySegment allocatePinned(long bytes)
    {
        return AircompressCuda.hostAlloc(bytes);
    }

    public static void freePinned(MemorySegment segment)
    {
        AircompressCuda.hostFree(segment);
    }

    /**
     * @param dst may be null for XXH64 (outLength[i] then receives the 64-bit hash, seed 0)
     * @throws IllegalStateException when the batch as a whole could not run (CUDA error, bad arguments); per-block problems are
     * reported through Result.status
     */
    public Result run(int op, MemorySegment src, long[] srcOffset, long[] srcLength, MemorySegment dst, long[] dstOffset, long[] dstCapacity)
    {
        int n = srcOffset.length;
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment srcOff = arena.allocateFrom(ValueLayout.JAVA_LONG, srcOffset);
            MemorySegment srcLen = arena.allocateFrom(ValueLayout.JAVA_LONG, srcLength);
            MemorySegment dstOff = dst == null ? MemorySegment.NULL : arena.allocateFrom(ValueLayout.JAVA_LONG, dstOffset);
            MemorySegment dstCap = dst == null ? MemorySegment.NULL : arena.allocateFrom(ValueLayout.JAVA_LONG, dstCapacity);
            MemorySegment outLen = arena.allocate(ValueLayout.JAVA_LONG, n);
            MemorySegment status = arena.allocate(ValueLayout.JAVA_INT, n);
            int result = AircompressCuda.batch(context.handle(), op, src, srcOff, srcLen, dst == null ? MemorySegment.NULL : dst,
                    dstOff, dstCap, outLen, status, n, 0, 0);
            if (result != 0) {
                throw new IllegalStateException("acc_batch failed", AircompressCuda.toException(-result, 0));
            }
            return new Result(outLen.toArray(ValueLayout.JAVA_LONG), status.toArray(ValueLayout.JAVA_INT));
        }
    }

    /** the exception a single-block call would have thrown for this status word and offset */
    public static RuntimeException toException(int status, long offset)
    {
        return AircompressCuda.toException(status, offset);
    }

    @Override
    public void close()
    {
        context.close();
    }
}
