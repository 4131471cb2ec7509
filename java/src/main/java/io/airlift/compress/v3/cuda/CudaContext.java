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
