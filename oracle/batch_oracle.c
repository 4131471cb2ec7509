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
            default: r = ORC_STATUS(ORC_E_ARGUMENT, 0);
        }
        out_len[i] = r;
        if (r < 0) failures++;
    }
    return failures;
}
