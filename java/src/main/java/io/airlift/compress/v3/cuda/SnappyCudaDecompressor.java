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

import io.airlift.compress.v3.Decompressor;
import io.airlift.compress.v3.MalformedInputException;

import java.lang.foreign.MemorySegment;

import static java.lang.Math.toIntExact;
import static java.lang.String.format;

/**
 * Snappy raw-format decompressor on the GPU: accepts and rejects exactly what SnappyJavaDecompressor does.
 * Implements the plain {@link Decompressor} interface (the codec-specific interfaces of the reference are sealed), one CUDA
 * context per instance; not thread-safe, like the reference's codec objects.
 */
public final class SnappyCudaDecompressor
        implements Decompressor, AutoCloseable
{
    private final CudaContext context;

    public SnappyCudaDecompressor()
    {
        this(0);
    }

    public SnappyCudaDecompressor(int device)
    {
        this.context = new CudaContext(device);
    }

    public static boolean isEnabled()
    {
        return AircompressCuda.isEnabled();
    }

    @Override
    public int decompress(byte[] input, int inputOffset, int inputLength, byte[] output, int outputOffset, int maxOutputLength)
            throws MalformedInputException
    {
        verifyRange(input, inputOffset, inputLength);
        verifyRange(output, outputOffset, maxOutputLength);
        return context.call(AircompressCuda.OP_SNAPPY_DECOMPRESS,
                MemorySegment.ofArray(input).asSlice(inputOffset, inputLength), inputLength,
                MemorySegment.ofArray(output).asSlice(outputOffset, maxOutputLength), maxOutputLength);
    }

    @Override
    public int decompress(MemorySegment input, MemorySegment output)
            throws MalformedInputException
    {
        return context.call(AircompressCuda.OP_SNAPPY_DECOMPRESS, input, input.byteSize(), output, output.byteSize());
    }

    /** SnappyDecompressor.getUncompressedLength: the length recorded in the stream's preamble */
    public int getUncompressedLength(byte[] compressed, int compressedOffset)
    {
        verifyRange(compressed, compressedOffset, compressed.length - compressedOffset);
        int length = compressed.length - compressedOffset;
        return toIntExact(AircompressCuda.snappyUncompressedLength(MemorySegment.ofArray(compressed).asSlice(compressedOffset, length), length));
    }

    private static void verifyRange(byte[] data, int offset, int length)
    {
        java.util.Objects.requireNonNull(data, "data is null");
        if (offset < 0 || length < 0 || offset + length > data.length) {
            throw new IllegalArgumentException(format("Invalid offset or length (%s, %s) in array of length %s", offset, length, data.length));
        }
    }

    @Override
    public void close()
    {
        context.close();
    }
}
