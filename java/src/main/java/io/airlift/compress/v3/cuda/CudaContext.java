package io.airlift.compress.v3.cuda;

import java.lang.foreign.MemorySegment;
import java.lang.ref.Cleaner;

/**
 * One acc_ctx (CUDA stream, staging buffers, work counters) per codec object.  Like the reference's codec objects it is
 * not thread-safe: one owner thread at a time; different contexts are fully concurrent.
 */
final class CudaContext
        implements AutoCloseable
{
    private static final Cleaner CLEANER = Cleaner.create();

    private final MemorySegment handle;
    private final Cleaner.Cleanable cleanable;

    CudaContext(int device)
    {
        AircompressCuda.verifyEnabled();
        MemorySegment ctx = AircompressCuda.init(device);
        this.handle = ctx;
        this.cleanable = CLEANER.register(this, () -> AircompressCuda.destroy(ctx));
    }

    MemorySegment handle()
    {
        return handle;
    }

    /** runs one single-block call and translates a failure into the exception the Java codec would throw */
    int call(int op, MemorySegment input, long inputLength, MemorySegment output, long maxOutputLength)
    {
        long result = AircompressCuda.call(op, handle, input, inputLength, output, maxOutputLength);
        if (result >= 0) {
            return Math.toIntExact(result);
        }
        long[] offset = new long[1];
        int status = AircompressCuda.lastError(handle, offset);
        if (op == AircompressCuda.OP_LZ4_DECOMPRESS && (status >>> 8) == AircompressCuda.R_LZ4_ZERO_CAPACITY) {
            return -1;   // Lz4RawDecompressor returns -1 for a zero-capacity output, it does not throw
        }
        throw AircompressCuda.toException(status, offset[0]);
    }

    @Override
    public void close()
    {
        cleanable.clean();
    }
}
