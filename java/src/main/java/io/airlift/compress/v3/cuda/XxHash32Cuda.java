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
 * One-shot XXH32 on the GPU: the counterpart of the static XxHash32Hasher.hash(...) overloads
 * (xxhash/XxHash32Hasher.java:18-50; the Java implementation is XxHash32JavaHasher.hash, :68-109).  XxHash32Hasher is
 * sealed, so this is a stand-alone class; the streaming object (update / digest) is not part of this path.  The LZ4 frame
 * format is its user (header, block and content checksums); for the block checksums of a whole frame use
 * {@link CudaBatch} with OP_XXH32: one launch hashes all blocks.
 */
public final class XxHash32Cuda
        implements AutoCloseable
{
    public static final int DEFAULT_SEED = 0;

    private final CudaContext context;

    public XxHash32Cuda()
    {
        this(0);
    }

    public XxHash32Cuda(int device)
    {
        this.context = new CudaContext(device);
    }

    public static boolean isEnabled()
    {
        return AircompressCuda.isEnabled();
    }

    public int hash(byte[] input)
    {
        return hash(input, 0, input.length, DEFAULT_SEED);
    }

    public int hash(byte[] input, int seed)
    {
        return hash(input, 0, input.length, seed);
    }

    public int hash(byte[] input, int offset, int length)
    {
        return hash(input, offset, length, DEFAULT_SEED);
    }

    public int hash(byte[] input, int offset, int length, int seed)
    {
        java.util.Objects.requireNonNull(input, "input is null");
        if (offset < 0 || length < 0 || offset + length > input.length) {
            throw new IllegalArgumentException(format("Invalid offset or length (%s, %s) in array of length %s", offset, length, input.length));
        }
        return AircompressCuda.xxh32(context.handle(), MemorySegment.ofArray(input).asSlice(offset, length), length, seed);
    }

    public int hash(MemorySegment input)
    {
        return hash(input, DEFAULT_SEED);
    }

    public int hash(MemorySegment input, int seed)
    {
        return AircompressCuda.xxh32(context.handle(), input, input.byteSize(), seed);
    }

    @Override
    public void close()
    {
        context.close();
    }
}
