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
package io.airlift.compress.v3.lz4;

import io.airlift.compress.v3.MalformedInputException;
import io.airlift.compress.v3.cuda.Lz4CudaDecompressor;

import java.lang.foreign.MemorySegment;

/**
 * LZ4 frame decompressor whose blocks are decoded on the GPU (see {@link Lz4FrameCudaCompressor} for where it lives and
 * why).  Accepts and rejects exactly what Lz4FrameJavaDecompressor does: the framing, the XXH32 checks and every message
 * are Lz4FrameCompression's; the block decoder reports malformed blocks with the reference's texts and offsets.
 * The batched form (header walk, ONE acc_batch for all blocks of the call, ONE XXH32 batch for the block checksums, then the
 * checks in the order of the sequential loop) is aircompressor_b200/lz4_frame.py in the CUDA repository.
 */
public final class Lz4FrameCudaDecompressor
        implements Lz4FrameDecompressor, AutoCloseable
{
    private final Lz4CudaDecompressor blockDecompressor;

    public Lz4FrameCudaDecompressor()
    {
        this(0);
    }

    public Lz4FrameCudaDecompressor(int device)
    {
        this.blockDecompressor = new Lz4CudaDecompressor(device);
    }

    public static boolean isEnabled()
    {
        return Lz4CudaDecompressor.isEnabled();
    }

    @Override
    public int decompress(byte[] input, int inputOffset, int inputLength, byte[] output, int outputOffset, int maxOutputLength)
            throws MalformedInputException
    {
        return Lz4FrameCompression.decompress(blockDecompressor, input, inputOffset, inputLength, output, outputOffset, maxOutputLength);
    }

    @Override
    public int decompress(MemorySegment input, MemorySegment output)
            throws MalformedInputException
    {
        return Lz4FrameCompression.decompress(blockDecompressor, input, output);
    }

    @Override
    public void close()
    {
        blockDecompressor.close();
    }
}
