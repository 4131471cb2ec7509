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
/*
 * aircompress-cuda: Java binding of libaircompress_cuda.so (see INTEGRATION.md).
 *
 * This file is an addition to airlift/aircompressor (package io.airlift.compress.v3.cuda) and uses the project's own
 * FFM loader (io.airlift.compress.v3.internal.NativeLoader / NativeSignature).  It has NOT been compiled in the
 * repository that carries it (no JDK in that build image); the C ABI it binds is exercised there through ctypes.
 */
package io.airlift.compress.v3.cuda;

import io.airlift.compress.v3.MalformedInputException;
import io.airlift.compress.v3.internal.NativeLoader.Symbols;
import io.airlift.compress.v3.internal.NativeSignature;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;
import java.util.Optional;

import static io.airlift.compress.v3.internal.NativeLoader.loadSymbols;
import static java.lang.invoke.MethodHandles.lookup;

/**
 * Static downcall handles for the C ABI declared in include/aircompress_cuda.h.  Every export uses only
 * int / long / MemorySegment, which is all NativeLoader supports.
 */
final class AircompressCuda
{
    // status word = code | reason << 8 (aircompress_cuda.h)
    static final int E_MALFORMED = 1;
    static final int E_DST_TOO_SMALL = 2;
    static final int E_ARGUMENT = 3;
    static final int E_CUDA = 4;
    static final int E_UNSUPPORTED = 5;
    static final int R_LZ4_ZERO_CAPACITY = 6;

    // batch op codes
    static final int OP_LZ4_COMPRESS = 0;
    static final int OP_LZ4_DECOMPRESS = 1;
    static final int OP_SNAPPY_COMPRESS = 2;
    static final int OP_SNAPPY_DECOMPRESS = 3;
    static final int OP_ZSTD_COMPRESS = 4;
    static final int OP_ZSTD_DECOMPRESS = 5;
    static final int OP_XXH64 = 6;
    static final int OP_XXH32 = 7;

    private record MethodHandles(
            @NativeSignature(name = "acc_device_count", returnType = int.class, argumentTypes = {}) MethodHandle deviceCount,
            @NativeSignature(name = "acc_init", returnType = MemorySegment.class, argumentTypes = int.class) MethodHandle init,
            @NativeSignature(name = "acc_init_error", returnType = int.class, argumentTypes = {}) MethodHandle initError,
            @NativeSignature(name = "acc_destroy", returnType = void.class, argumentTypes = MemorySegment.class) MethodHandle destroy,
            @NativeSignature(name = "acc_host_alloc", returnType = MemorySegment.class, argumentTypes = long.class) MethodHandle hostAlloc,
            @NativeSignature(name = "acc_host_free", returnType = void.class, argumentTypes = MemorySegment.class) MethodHandle hostFree,
            @NativeSignature(name = "acc_last_error", returnType = int.class, argumentTypes = {MemorySegment.class, MemorySegment.class}) MethodHandle lastError,
            @NativeSignature(name = "acc_code_name", returnType = MemorySegment.class, argumentTypes = int.class) MethodHandle codeName,
            @NativeSignature(name = "acc_reason_text", returnType = MemorySegment.class, argumentTypes = int.class) MethodHandle reasonText,
            @NativeSignature(name = "acc_lz4_compress_bound", returnType = long.class, argumentTypes = long.class) MethodHandle lz4Bound,
            @NativeSignature(name = "acc_snappy_compress_bound", returnType = long.class, argumentTypes = long.class) MethodHandle snappyBound,
            @NativeSignature(name = "acc_zstd_compress_bound", returnType = long.class, argumentTypes = long.class) MethodHandle zstdBound,
            @NativeSignature(name = "acc_lz4_compress", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, MemorySegment.class, long.class}) MethodHandle lz4Compress,
            @NativeSignature(name = "acc_lz4_decompress", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, MemorySegment.class, long.class}) MethodHandle lz4Decompress,
            @NativeSignature(name = "acc_snappy_compress", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, MemorySegment.class, long.class}) MethodHandle snappyCompress,
            @NativeSignature(name = "acc_snappy_decompress", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, MemorySegment.class, long.class}) MethodHandle snappyDecompress,
            @NativeSignature(name = "acc_zstd_compress", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, MemorySegment.class, long.class}) MethodHandle zstdCompress,
            @NativeSignature(name = "acc_zstd_decompress", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, MemorySegment.class, long.class}) MethodHandle zstdDecompress,
            @NativeSignature(name = "acc_snappy_uncompressed_length", returnType = long.class, argumentTypes = {MemorySegment.class, long.class, MemorySegment.class}) MethodHandle snappyUncompressedLength,
            @NativeSignature(name = "acc_zstd_frame_content_size", returnType = long.class, argumentTypes = {MemorySegment.class, long.class, MemorySegment.class}) MethodHandle zstdFrameContentSize,
            @NativeSignature(name = "acc_xxh64", returnType = long.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, long.class}) MethodHandle xxh64,
            @NativeSignature(name = "acc_xxh32", returnType = int.class, argumentTypes = {MemorySegment.class, MemorySegment.class, long.class, int.class}) MethodHandle xxh32,
            @NativeSignature(name = "acc_batch", returnType = int.class, argumentTypes = {MemorySegment.class, int.class,
                    MemorySegment.class, MemorySegment.class, MemorySegment.class, MemorySegment.class, MemorySegment.class, MemorySegment.class,
                    MemorySegment.class, MemorySegment.class, long.class, int.class, long.class}) MethodHandle batch) {}

    private static final Optional<LinkageError> LINKAGE_ERROR;
    private static final MethodHandles H;

    static {
        Symbols<MethodHandles> symbols = loadSymbols("aircompress_cuda", MethodHandles.class, lookup());
        LINKAGE_ERROR = symbols.linkageError();
        H = symbols.symbols();
    }

    private AircompressCuda() {}

    /** true when the library loaded AND a CUDA device is visible */
    static boolean isEnabled()
    {
        if (LINKAGE_ERROR.isPresent()) {
            return false;
        }
        try {
            return (int) H.deviceCount().invokeExact() > 0;
        }
        catch (Throwable e) {
            return false;
        }
    }

    static void verifyEnabled()
    {
        if (LINKAGE_ERROR.isPresent()) {
            throw new IllegalStateException("aircompress_cuda native library is not enabled", LINKAGE_ERROR.get());
        }
        if (!isEnabled()) {
            throw new IllegalStateException("aircompress_cuda: no CUDA device available");
        }
    }

    // ---- context ----

    static MemorySegment init(int device)
    {
        try {
            MemorySegment ctx = (MemorySegment) H.init().invokeExact(device);
            if (ctx.equals(MemorySegment.NULL)) {
                int status = (int) H.initError().invokeExact();
                throw new IllegalStateException("acc_init failed: " + describe(status));
            }
            return ctx;
        }
        catch (Error | RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    static void destroy(MemorySegment ctx)
    {
        try {
            H.destroy().invokeExact(ctx);
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    // ---- single-block calls: return bytes written, or throw exactly what the Java codecs throw ----

    static long bound(int op, long n)
    {
        try {
            return switch (op) {
                case OP_LZ4_COMPRESS -> (long) H.lz4Bound().invokeExact(n);
                case OP_SNAPPY_COMPRESS -> (long) H.snappyBound().invokeExact(n);
                case OP_ZSTD_COMPRESS -> (long) H.zstdBound().invokeExact(n);
                default -> throw new IllegalArgumentException("not a compress op: " + op);
            };
        }
        catch (Error | RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    /** @return bytes written (>= 0), or the negated status word (< 0) -- see {@link #lastErrorOffset} */
    static long call(int op, MemorySegment ctx, MemorySegment input, long inputLength, MemorySegment output, long maxOutputLength)
    {
        try {
            return switch (op) {
                case OP_LZ4_COMPRESS -> (long) H.lz4Compress().invokeExact(ctx, input, inputLength, output, maxOutputLength);
                case OP_LZ4_DECOMPRESS -> (long) H.lz4Decompress().invokeExact(ctx, input, inputLength, output, maxOutputLength);
                case OP_SNAPPY_COMPRESS -> (long) H.snappyCompress().invokeExact(ctx, input, inputLength, output, maxOutputLength);
                case OP_SNAPPY_DECOMPRESS -> (long) H.snappyDecompress().invokeExact(ctx, input, inputLength, output, maxOutputLength);
                case OP_ZSTD_COMPRESS -> (long) H.zstdCompress().invokeExact(ctx, input, inputLength, output, maxOutputLength);
                case OP_ZSTD_DECOMPRESS -> (long) H.zstdDecompress().invokeExact(ctx, input, inputLength, output, maxOutputLength);
                default -> throw new IllegalArgumentException("unknown op: " + op);
            };
        }
        catch (Error | RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    static long xxh64(MemorySegment ctx, MemorySegment input, long length, long seed)
    {
        try {
            return (long) H.xxh64().invokeExact(ctx, input, length, seed);
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    static int xxh32(MemorySegment ctx, MemorySegment input, long length, int seed)
    {
        try {
            return (int) H.xxh32().invokeExact(ctx, input, length, seed);
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    /** status word of the last failed single-block call of this context; the offset lands in offsetOut[0] */
    static int lastError(MemorySegment ctx, long[] offsetOut)
    {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment offset = arena.allocate(ValueLayout.JAVA_LONG);
            int status = (int) H.lastError().invokeExact(ctx, offset);
            offsetOut[0] = offset.get(ValueLayout.JAVA_LONG, 0);
            return status;
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    static long snappyUncompressedLength(MemorySegment input, long length)
    {
        return headerQuery(H.snappyUncompressedLength(), input, length, false);
    }

    /** -1 when the frame does not record its content size (like FrameHeader.contentSize in the Java code) */
    static long zstdFrameContentSize(MemorySegment input, long length)
    {
        return headerQuery(H.zstdFrameContentSize(), input, length, true);
    }

    private static long headerQuery(MethodHandle handle, MemorySegment input, long length, boolean minusOneIsAnswer)
    {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment offset = arena.allocate(ValueLayout.JAVA_LONG);
            long result = (long) handle.invokeExact(input, length, offset);
            if (result >= 0 || (minusOneIsAnswer && result == -1)) {
                return result;
            }
            throw toException((int) -result, offset.get(ValueLayout.JAVA_LONG, 0));
        }
        catch (Error | RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    // ---- batches: all segments must be native (ideally pinned, from hostAlloc) ----

    static int batch(MemorySegment ctx, int op, MemorySegment srcBase, MemorySegment srcOff, MemorySegment srcLen,
            MemorySegment dstBase, MemorySegment dstOff, MemorySegment dstCap, MemorySegment outLen, MemorySegment status, long n, int flags, long stream)
    {
        try {
            return (int) H.batch().invokeExact(ctx, op, srcBase, srcOff, srcLen, dstBase, dstOff, dstCap, outLen, status, n, flags, stream);
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    static MemorySegment hostAlloc(long bytes)
    {
        try {
            MemorySegment p = (MemorySegment) H.hostAlloc().invokeExact(bytes);
            if (p.equals(MemorySegment.NULL)) {
                throw new OutOfMemoryError("acc_host_alloc(" + bytes + ")");
            }
            return p.reinterpret(bytes);
        }
        catch (Error | RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    static void hostFree(MemorySegment p)
    {
        try {
            H.hostFree().invokeExact(p);
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }

    // ---- error translation: the same exception types and messages as the Java codecs ----

    static RuntimeException toException(int status, long offset)
    {
        int code = status & 0xFF;
        int reason = status >>> 8;
        String text = cString(H.reasonText(), reason);
        return switch (code) {
            case E_MALFORMED -> new MalformedInputException(offset, text);
            case E_DST_TOO_SMALL -> new IllegalArgumentException("Output buffer too small: " + text);
            case E_ARGUMENT -> new IllegalArgumentException(text);
            case E_UNSUPPORTED -> new UnsupportedOperationException("unsupported by this build: " + text);
            default -> new IllegalStateException(describe(status));
        };
    }

    private static String describe(int status)
    {
        return cString(H.codeName(), status & 0xFF) + " (reason " + (status >>> 8) + ")";
    }

    private static String cString(MethodHandle handle, int argument)
    {
        try {
            MemorySegment p = (MemorySegment) handle.invokeExact(argument);
            return p.equals(MemorySegment.NULL) ? "" : p.reinterpret(Long.MAX_VALUE).getString(0);
        }
        catch (Throwable e) {
            throw new AssertionError("should not reach here", e);
        }
    }
}
