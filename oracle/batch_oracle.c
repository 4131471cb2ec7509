/*
 * batch_oracle.c -- OpenMP driver that runs the oracle one independent block per task
 * (TEST INFRASTRUCTURE, see oracle.h).  Mirrors how the reference is driven: one
 * Compressor/Decompressor call per block (AbstractTestCompression.java:362-393,
 * benchmark/CompressionBenchmark.java:103-117), no state carried between calls.
 */
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

int32_t orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int64_t orc_batch(int32_t op, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                  uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                  int64_t *out_len, int64_t n, int32_t threads)
{
    int64_t failures = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) reduction(+ : failures)
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *s = src_base + src_off[i];
        uint8_t *d = dst_base ? dst_base + (dst_off ? dst_off[i] : 0) : 0;
        int64_t cap = dst_cap ? dst_cap[i] : 0;
        int64_t r;
        switch (op) {
            case 0: r = orc_lz4_compress(s, src_len[i], d, cap); break;
            case 1: r = orc_lz4_decompress(s, src_len[i], d, cap, 0); break;
            case 2: r = orc_snappy_compress(s, src_len[i], d, cap); break;
            case 3: r = orc_snappy_decompress(s, src_len[i], d, cap, 0); break;
            case 4: r = orc_zstd_compress(s, src_len[i], d, cap); break;
            case 5: r = orc_zstd_decompress(s, src_len[i], d, cap, 0); break;
            case 6: r = (int64_t) orc_xxh64(s, src_len[i], 0); out_len[i] = r; continue;
            case 7: r = (int64_t) orc_xxh32(s, src_len[i], 0); out_len[i] = r; continue;
            default: r = ORC_STATUS(ORC_E_ARGUMENT, 0);
        }
        out_len[i] = r;
        if (r < 0) failures++;
    }
    return failures;
}

/*
 * The same per-block loop over the reference's bundled NATIVE libraries (what Lz4Native / SnappyNative / ZstdNative bind,
 * lz4/Lz4Native.java:97-146, snappy/SnappyNative.java:37-56, zstd/ZstdNative.java:108-163): `fn` is the address of
 * LZ4_compress_fast / LZ4_decompress_safe / snappy_compress / snappy_uncompress / ZSTD_compress / ZSTD_decompress,
 * resolved by the caller with dlopen (oracle/pyoracle.py RefNative).  Used by bench.py to report the reference's native
 * CPU path next to the port of its Java path.  Returns the number of failed blocks.
 */
int64_t orc_native_batch(int32_t op, void *fn, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                         uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *out_len, int64_t n, int32_t threads)
{
    typedef int (*lz4c_t)(const char *, char *, int, int, int);
    typedef int (*lz4d_t)(const char *, char *, int, int);
    typedef int (*snp_t)(const char *, size_t, char *, size_t *);
    typedef size_t (*zc_t)(void *, size_t, const void *, size_t, int);
    typedef size_t (*zd_t)(void *, size_t, const void *, size_t);
    int64_t failures = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) reduction(+ : failures)
    for (int64_t i = 0; i < n; i++) {
        const char *s = (const char *) (src_base + src_off[i]);
        char *d = (char *) (dst_base + dst_off[i]);
        const int64_t cap = dst_cap[i];
        int64_t r = -1;
        switch (op) {
            case 0: r = ((lz4c_t) fn)(s, d, (int) src_len[i], (int) cap, 1); if (r <= 0) r = -1; break;
            case 1: r = ((lz4d_t) fn)(s, d, (int) src_len[i], (int) cap); break;
            case 2: case 3: { size_t len = (size_t) cap; int st = ((snp_t) fn)(s, (size_t) src_len[i], d, &len); r = st == 0 ? (int64_t) len : -1; break; }
            case 4: { size_t z = ((zc_t) fn)(d, (size_t) cap, s, (size_t) src_len[i], 3); r = z <= (size_t) cap ? (int64_t) z : -1; break; }
            case 5: { size_t z = ((zd_t) fn)(d, (size_t) cap, s, (size_t) src_len[i]); r = z <= (size_t) cap ? (int64_t) z : -1; break; }
            default: break;
        }
        out_len[i] = r;
        if (r < 0) failures++;
    }
    return failures;
}
