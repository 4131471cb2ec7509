/*
 * aircompress_cuda.h -- C ABI of libaircompress_cuda.so, the B200 (sm_100a) batched block-compression
 * engine that sits behind io.airlift.compress.v3.Compressor / Decompressor.
 *
 * This is exactly what the reference's FFM layer would bind for this path.  The reference binds its
 * native codecs through records of MethodHandles annotated with @NativeSignature and resolved by
 * NativeLoader.loadSymbols (internal/NativeLoader.java:66-117); only byte/int/long/MemorySegment are
 * legal argument/return types there (internal/NativeLoader.java:119-153).  Therefore every function
 * below uses only int8/int32/int64 values and pointers -- no structs by value, no size_t.
 *
 * Each entry point names the reference binding it replaces (paths relative to
 * /root/reference/src/main/java/io/airlift/compress/v3/).  INTEGRATION.md shows the Java side.
 *
 * Conventions
 *   - Single-block functions return int64: >= 0 bytes written, < 0 = -(status) where
 *     status = code | reason << 8 (ACC_E_* / ACC_R_* below).  acc_last_error() returns the same status
 *     and the input offset the reference would pass to MalformedInputException(offset, reason).
 *   - Batch functions process n independent blocks (one Compressor/Decompressor call each in the
 *     reference) in one launch.  Per block i:  status[i] = 0 or the status word; out_len[i] = bytes
 *     written, or, when status[i] != 0, the error offset.  The batch return value is 0 when the launch
 *     happened (inspect status[]), or -(ACC_E_CUDA...) when it could not.
 *   - No CPU fallback exists.  Without a usable GPU acc_init returns NULL and acc_init_error() says why.
 *   - One acc_ctx = one owner thread at a time (the reference's codec objects are not thread-safe,
 *     lz4/Lz4JavaCompressor.java:27-29); distinct contexts are fully concurrent.
 */
#ifndef AIRCOMPRESS_CUDA_H
#define AIRCOMPRESS_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (low 8 bits) ---------------------------------------------------------------- */
#define ACC_OK               0
#define ACC_E_MALFORMED      1  /* -> MalformedInputException(offset, reason)  (MalformedInputException.java:21-35) */
#define ACC_E_DST_TOO_SMALL  2  /* -> IllegalArgumentException "Output buffer too small" family */
#define ACC_E_ARGUMENT       3  /* -> IllegalArgumentException (bad lengths / bounds) */
#define ACC_E_CUDA           4  /* CUDA runtime failure (acc_init_error / acc_last_error reason = cudaError_t) */
#define ACC_E_UNSUPPORTED    5  /* input is valid but outside what this build handles (never a silent fallback) */

/* ---- reasons (status >> 8); text via acc_reason_text() ---------------------------------------- */
#define ACC_R_NONE                  0   /* "Malformed input" */
#define ACC_R_INPUT_EMPTY           1   /* lz4/Lz4RawDecompressor.java:48-50 */
#define ACC_R_LAST_LITERAL_OUTSIDE  2   /* lz4/Lz4RawDecompressor.java:84-86 */
#define ACC_R_ALL_INPUT_CONSUMED    3   /* lz4/Lz4RawDecompressor.java:88-90 */
#define ACC_R_OFFSET_OUTSIDE        4   /* lz4/Lz4RawDecompressor.java:116-119 */
#define ACC_R_LAST5_LITERALS        5   /* lz4/Lz4RawDecompressor.java:168-171 */
#define ACC_R_LZ4_ZERO_CAPACITY     6   /* lz4/Lz4RawDecompressor.java:52-57 (Java returns -1) */
#define ACC_R_SNAPPY_TRUNCATED      7   /* snappy/SnappyRawDecompressor.java:315-320 */
#define ACC_R_SNAPPY_VARINT_HIGHBIT 8   /* snappy/SnappyRawDecompressor.java:303-305 */
#define ACC_R_SNAPPY_NEG_LENGTH     9   /* snappy/SnappyRawDecompressor.java:309-311 */
#define ACC_R_SNAPPY_LEN_GT_CAP     10  /* snappy/SnappyRawDecompressor.java:49-50 */
#define ACC_R_SNAPPY_LEN_MISMATCH   11  /* snappy/SnappyRawDecompressor.java:61-65 */
#define ACC_R_MAX_OUTPUT_TOO_SMALL  12  /* lz4/Lz4RawCompressor.java:87-89, snappy/SnappyRawCompressor.java:87-90, zstd/ZstdFrameCompressor.java */
#define ACC_R_MAX_INPUT_EXCEEDED    13  /* lz4/Lz4RawCompressor.java:83-85 */
#define ACC_R_ZSTD_BASE             32  /* zstd reasons: see acc_reason_text() */

/* ---- flags for batch calls --------------------------------------------------------------------- */
#define ACC_F_DEVICE_POINTERS 1  /* src/dst bases and all index/result arrays are device pointers; the call is
                                    asynchronous on `stream` (no host<->device copies, no synchronisation) */

typedef struct acc_ctx acc_ctx;

/* ---- lifecycle (replaces the static initialisers + isEnabled()/verifyEnabled() of
 *      lz4/Lz4Native.java:42-85, snappy/SnappyNative.java, zstd/ZstdNative.java) ------------------ */
int32_t  acc_device_count(void);                 /* 0 when no driver / no GPU; never crashes at load time */
acc_ctx *acc_init(int32_t device);               /* NULL on failure, see acc_init_error() */
int32_t  acc_init_error(void);                   /* status word of the last failed acc_init on this thread */
void     acc_destroy(acc_ctx *ctx);
void    *acc_host_alloc(int64_t bytes);          /* pinned host memory for batch staging */
void     acc_host_free(void *p);
int32_t  acc_last_error(acc_ctx *ctx, int64_t *offset);  /* status word + offset of the last failed single-block call */
const char *acc_code_name(int32_t code);
const char *acc_reason_text(int32_t reason);     /* the reference's exception message for that reason */
int32_t  acc_device_numa_node(int32_t device);   /* NUMA node of the GPU's PCIe root (sysfs), -1 when unknown */
int32_t  acc_bind_host_thread(int32_t device);   /* pins the CALLING THREAD to that node's CPUs (call before allocating / pinning the
                                                    buffers of host-pointer batches: pages are placed by first touch); returns the
                                                    node, or -1 and changes nothing.  One host thread + one context per device is the
                                                    multi-GPU model */
int32_t  acc_sm_count(acc_ctx *ctx);
int64_t  acc_kernel_launches(acc_ctx *ctx);      /* kernels launched through this ctx so far (bench.py gpu_launches) */

/* per-context counters since acc_init (the reference has no tracing; this is the hook a caller can poll).
 * Copies min(words, ACC_STATS_WORDS) int64 values into out[] and returns that count. */
#define ACC_STAT_BATCHES        0   /* kernel batches enqueued (host- and device-pointer calls; pipelined runs count each) */
#define ACC_STAT_BLOCKS         1   /* blocks in those batches */
#define ACC_STAT_LAUNCHES       2   /* kernel launches */
#define ACC_STAT_HOST_CALLS     3   /* host-pointer calls (single-block calls included) */
#define ACC_STAT_H2D_BYTES      4   /* bytes uploaded by host-pointer calls (payload + index arrays) */
#define ACC_STAT_D2H_BYTES      5   /* bytes downloaded by host-pointer calls (output windows + result arrays) */
#define ACC_STAT_LAST_CALL_US   6   /* wall time of the last host-pointer call: staging, upload, kernels, download, sync */
#define ACC_STAT_TOTAL_CALL_US  7   /* sum of the above over all host-pointer calls */
#define ACC_STATS_WORDS         8
int32_t  acc_get_stats(acc_ctx *ctx, int64_t *out, int32_t words);

/* ---- bounds: replace LZ4_compressBound (lz4/Lz4Native.java:31), snappy_max_compressed_length
 *      (snappy/SnappyNative.java:70), ZSTD_compressBound (zstd/ZstdNative.java:29).  Values follow the
 *      Java compressors' maxCompressedLength (lz4/Lz4RawCompressor.java:64-67,
 *      snappy/SnappyRawCompressor.java:47-70, zstd/ZstdJavaCompressor.java:31-40). ------------------ */
int64_t acc_lz4_compress_bound(int64_t n);
int64_t acc_snappy_compress_bound(int64_t n);
int64_t acc_zstd_compress_bound(int64_t n);

/* ---- single block, host pointers (what Compressor.compress(byte[]...) / decompress(...) reach):
 *      LZ4_compress_fast_extState / LZ4_decompress_safe   (lz4/Lz4Native.java:34-38,97-146)
 *      snappy_compress / snappy_uncompress                 (snappy/SnappyNative.java:68-75,105-138)
 *      ZSTD_compress / ZSTD_decompress                     (zstd/ZstdNative.java:30-33,108-143)
 *      The pointers may be pageable Java-heap memory pinned for the call (FFM critical downcall);
 *      the library stages through its own pinned buffers and never retains them. ------------------- */
int64_t acc_lz4_compress(acc_ctx *ctx, const void *src, int64_t src_len, void *dst, int64_t dst_cap);
int64_t acc_lz4_decompress(acc_ctx *ctx, const void *src, int64_t src_len, void *dst, int64_t dst_cap);
int64_t acc_snappy_compress(acc_ctx *ctx, const void *src, int64_t src_len, void *dst, int64_t dst_cap);
int64_t acc_snappy_decompress(acc_ctx *ctx, const void *src, int64_t src_len, void *dst, int64_t dst_cap);
int64_t acc_zstd_compress(acc_ctx *ctx, const void *src, int64_t src_len, void *dst, int64_t dst_cap);
int64_t acc_zstd_decompress(acc_ctx *ctx, const void *src, int64_t src_len, void *dst, int64_t dst_cap);

/* header-only host helpers: snappy_uncompressed_length (snappy/SnappyNative.java:72-75; Java
 * SnappyRawDecompressor.getUncompressedLength :30-33) and ZSTD_getFrameContentSize
 * (zstd/ZstdNative.java:34; Java ZstdFrameDecompressor.getDecompressedSize :942-947).
 * Return >= 0 or -(status); *err_offset optional.  acc_zstd_frame_content_size returns -1 when the frame
 * does not record its size (error statuses of that function are always < -0x2000). */
int64_t acc_snappy_uncompressed_length(const void *src, int64_t src_len, int64_t *err_offset);
int64_t acc_zstd_frame_content_size(const void *src, int64_t src_len, int64_t *err_offset);

/* XXH64 one-shot: replaces XXH64(input, length, seed) (xxhash/XxHash64Bindings.java:32-34,81-100). */
int64_t acc_xxh64(acc_ctx *ctx, const void *src, int64_t len, int64_t seed);

/* XXH32 one-shot: replaces XXH32(input, length, seed) (xxhash/XxHash32Bindings.java; Java: XxHash32JavaHasher.hash,
 * xxhash/XxHash32JavaHasher.java:68-109) -- the checksum of the LZ4 frame format (lz4/Lz4FrameCompression.java:95,216,285,307). */
int32_t acc_xxh32(acc_ctx *ctx, const void *src, int64_t len, int32_t seed);

/* ---- batches of independent blocks (the GPU-shaped entry points; bound non-critical from Java).
 *      op codes select codec + direction; all share one signature so the Java record stays small. */
#define ACC_OP_LZ4_COMPRESS      0
#define ACC_OP_LZ4_DECOMPRESS    1
#define ACC_OP_SNAPPY_COMPRESS   2
#define ACC_OP_SNAPPY_DECOMPRESS 3
#define ACC_OP_ZSTD_COMPRESS     4
#define ACC_OP_ZSTD_DECOMPRESS   5
#define ACC_OP_XXH64             6   /* out_len[i] receives the 64-bit hash (seed 0); dst_* unused (may be NULL) */
#define ACC_OP_XXH32             7   /* out_len[i] receives the 32-bit hash, zero-extended (seed 0 through acc_batch) */

/*
 * block i reads  src_base[src_off[i] .. src_off[i]+src_len[i])  and writes at most dst_cap[i] bytes at
 * dst_base + dst_off[i].  `stream` is a CUstream/cudaStream_t handle (0 = the context's own non-blocking
 * stream; pass 1 (cudaStreamLegacy) or 2 (cudaStreamPerThread) to target CUDA's default streams);
 * with ACC_F_DEVICE_POINTERS the work is only enqueued (at most 100 batches may be in flight per context:
 * every launch takes one or two of the context's 256 work-stealing counters, which are reused round-robin).
 * Batches of ONE context are ordered: a batch enqueued on a different stream than the previous one first waits for
 * it (the context's scratch and counters are shared), so use one context per concurrent stream of work.
 * Device buffers: kernels read whole aligned words (4 bytes; 8 in the XXH64 kernel, 16 in the XXH32 kernel, 32 in the parse
 * kernel of the record path), i.e. a few bytes in front of / behind a block's [src_off, src_off + src_len) range -- never beyond the
 * 32-byte-aligned extent of src_base's allocation (cudaMalloc sizes are multiples of 256 bytes), so pad a sub-allocated
 * source buffer to a multiple of 32 bytes.
 * Without it the library copies host->device,
 * runs, copies results back and synchronises before returning; large batches are cut into runs of consecutive
 * blocks whose upload, kernel and download overlap (see acc_set_tuning key 3).
 * Output contract: bytes [0, out_len[i]) of window i are the result; the rest of the window
 * [out_len[i], dst_cap[i]) is unspecified after a host-pointer batch of more than one block (whole windows are
 * copied back so that touching windows merge into one transfer); bytes outside every window are never written.
 * The single-block entry points above write only the bytes they return, like the reference codecs.
 */
int32_t acc_batch(acc_ctx *ctx, int32_t op,
                  const void *src_base, const int64_t *src_off, const int64_t *src_len,
                  void *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                  int64_t *out_len, int32_t *status, int64_t n, int32_t flags, int64_t stream);

/* per-op aliases with the names a Java record component would carry (thin wrappers over acc_batch) */
int32_t acc_lz4_compress_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, void *, const int64_t *, const int64_t *, int64_t *, int32_t *, int64_t, int32_t, int64_t);
int32_t acc_lz4_decompress_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, void *, const int64_t *, const int64_t *, int64_t *, int32_t *, int64_t, int32_t, int64_t);
int32_t acc_snappy_compress_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, void *, const int64_t *, const int64_t *, int64_t *, int32_t *, int64_t, int32_t, int64_t);
int32_t acc_snappy_decompress_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, void *, const int64_t *, const int64_t *, int64_t *, int32_t *, int64_t, int32_t, int64_t);
int32_t acc_zstd_compress_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, void *, const int64_t *, const int64_t *, int64_t *, int32_t *, int64_t, int32_t, int64_t);
int32_t acc_zstd_decompress_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, void *, const int64_t *, const int64_t *, int64_t *, int32_t *, int64_t, int32_t, int64_t);
int32_t acc_xxh64_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, int64_t *, int64_t, int32_t, int64_t);
/* (ctx, src_base, src_off, src_len, hashes, n, seed, flags, stream): XXH32 of n buffers, hashes[i] zero-extended */
int32_t acc_xxh32_batch(acc_ctx *, const void *, const int64_t *, const int64_t *, int64_t *, int64_t, int32_t, int32_t, int64_t);

/* tuning knob used by bench.py sweeps: 0 restores the default. Returns the previous value.
 * key 0: resident CTAs per SM for the warp-per-block decode kernels;
 * key 1: LZ4 / Snappy decode path: 1 = the step decoder (one warp walks and copies a block), 2 = the record path (parse kernel +
 *        execute kernel, csrc/lz_records.cuh), 0 = automatic (the record path for Snappy batches of >= 49,152 blocks, where its
 *        fixed parse latency is paid back; the step decoder otherwise);
 * key 3: host-pointer batches, 1 = never split, k > 1 = split into k overlapped upload/kernel/download runs
 * (default: automatic, up to 16 runs of >= 4096 blocks and >= 32 MiB each); other keys are ignored. */
int32_t acc_set_tuning(acc_ctx *ctx, int32_t key, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* AIRCOMPRESS_CUDA_H */
