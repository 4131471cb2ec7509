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

import io.airlift.compress.v3.cuda.Lz4CudaCompressor;

import java.lang.foreign.MemorySegment;

/**
 * LZ4 frame compressor whose blocks are compressed on the GPU: the third implementation next to Lz4FrameJavaCompressor and
 * Lz4FrameNativeCompressor (lz4/Lz4FrameJavaCompressor.java:25-44).  It lives in this package because the shared framing,
 * Lz4FrameCompression, is package-private, and it needs two one-word edits in the reference: Lz4FrameCompressor's and
 * Lz4Compressor's permits lists (both interfaces are sealed).
 * <p>
 * This is the drop-in form: Lz4FrameCompression.compress walks the blocks of the frame and calls the block codec once per
 * block.  A frame is a list of independent blocks, so the GPU-shaped form hands all of them to acc_batch at once and then
 * assembles the frame -- aircompressor_b200/lz4_frame.py in the CUDA repository is that loop (same frames, same error
 * behaviour) and the model for a batched override of compress().
 */
public final class Lz4FrameCudaCompressor
        implements Lz4FrameCompressor, AutoCloseable
{
    private final Lz4CudaCompressor blockCompressor;

    public Lz4FrameCudaCompressor()
    {
        this(0);
    }

    public Lz4FrameCudaCompressor(int device)
    {
        this.blockCompressor = new Lz4CudaCompressor(device);
    }

    public static boolean isEnabled()
    {
        return Lz4CudaCompressor.isEnabled();
    }

    @Override
    public int maxCompressedLength(int uncompressedSize)
    {
        return Lz4FrameCompression.maxCompressedLength(uncompressedSize);
    }

    @Override
    public int compress(byte[] input, int inputOffset, int inputLength, byte[] output, int outputOffset, int maxOutputLength)
    {
        return Lz4FrameCompression.compress(blockCompressor, input, inputOffset, inputLength, output, outputOffset, maxOutputLength);
    }

    @Override
    public int compress(MemorySegment input, MemorySegment output)
    {
        return Lz4FrameCompression.compress(blockCompressor, input, output);
    }

    @Override
    public void close()
    {
        blockCompressor.close();
    }
}
