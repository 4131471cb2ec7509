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

import static java.lang.String.format;

/**
 * One-shot XXH64 on the GPU: the counterpart of the static XxHash64Hasher.hash(...) overloads
 * (xxhash/XxHash64Hasher.java:44-80).  XxHash64Hasher itself is sealed, so this is a stand-alone class; the streaming
 * object (update / digest) is not part of this path.  For many buffers use {@link CudaBatch} with OP_XXH64: one
 * kernel launch hashes the whole batch at HBM speed, whereas a single small buffer is dominated by the call overhead.
 */
public final class XxHash64Cuda
        implements AutoCloseable
{
    public static final long DEFAULT_SEED = 0;

    private final CudaContext context;

    public XxHash64Cuda()
    {
        this(0);
    }

    public XxHash64Cuda(int device)
    {
        this.context = new CudaContext(device);
    }

    public static boolean isEnabled()
    {
        return AircompressCuda.isEnabled();
    }

    public long hash(long value)
    {
        return hash(value, DEFAULT_SEED);
    }

    /** the 8 bytes of value in little-endian order, like XxHash64Hasher.hash(long, long) */
    public long hash(long value, long seed)
    {
        byte[] bytes = new byte[8];
        for (int i = 0; i < 8; i++) {
            bytes[i] = (byte) (value >>> (8 * i));
        }
        return hash(bytes, 0, 8, seed);
    }

    public long hash(byte[] input)
    {
        return hash(input, 0, input.length, DEFAULT_SEED);
    }

    public long hash(byte[] input, long seed)
    {
        return hash(input, 0, input.length, seed);
    }

    public long hash(byte[] input, int offset, int length)
    {
        return hash(input, offset, length, DEFAULT_SEED);
    }

    public long hash(byte[] input, int offset, int length, long seed)
    {
        java.util.Objects.requireNonNull(input, "input is null");
        if (offset < 0 || length < 0 || offset + length > input.length) {
            throw new IllegalArgumentException(format("Invalid offset or length (%s, %s) in array of length %s", offset, length, input.length));
        }
        return AircompressCuda.xxh64(context.handle(), MemorySegment.ofArray(input).asSlice(offset, length), length, seed);
    }

    public long hash(MemorySegment input, long seed)
    {
        return AircompressCuda.xxh64(context.handle(), input, input.byteSize(), seed);
    }

    @Override
    public void close()
    {
        context.close();
    }
}
