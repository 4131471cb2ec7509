/*
 * oracle.h -- CPU restatement ("port") of the reference's pure-Java codec kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this library,
 * and there only as the checker (or the timed CPU baseline), never as the thing shipped.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/main/java/io/airlift/compress/v3/).  Java semantics that matter are kept:
 * 32-bit int wrap-around, >>> as unsigned shift, long multiply wrap in hashes, byte & 0xFF loads,
 * little-endian unaligned loads.
 *
 * Parity pinning: see oracle/README.md -- golden vectors of the reference's own tests
 * (tests/test_oracle_golden.py) plus cross-decoding against the reference's bundled native
 * libraries (oracle/_ref) pin this restatement.
 *
 * Error convention (shared with include/aircompress_cuda.h so tests compare numbers directly):
 *   return >= 0 : bytes written;  return < 0 : -(status), status = code | reason << 8;
 *   *err_offset (optional) receives the offset the Java code passes to MalformedInputException.
 */
#ifndef AIRCOMPRESS_ORACLE_H
#define AIRCOMPRESS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* codes (low 8 bits of status) -- same numbering as include/aircompress_cuda.h */
enum {
    ORC_OK = 0,
    ORC_E_MALFORMED = 1,      /* MalformedInputException */
    ORC_E_DST_TOO_SMALL = 2,  /* IllegalArgumentException "Output buffer too small" & friends */
    ORC_E_ARGUMENT = 3        /* other IllegalArgumentException */
};

/* reasons (status >> 8) -- same numbering as include/aircompress_cuda.h */
enum {
    ORC_R_NONE = 0,                 /* "Malformed input" */
    ORC_R_INPUT_EMPTY = 1,          /* lz4: "input is empty" */
    ORC_R_LAST_LITERAL_OUTSIDE = 2, /* lz4: "attempt to write last literal outside of destination buffer" */
    ORC_R_ALL_INPUT_CONSUMED = 3,   /* lz4: "all input must be consumed" */
    ORC_R_OFFSET_OUTSIDE = 4,       /* lz4: "offset outside destination buffer" */
    ORC_R_LAST5_LITERALS = 5,       /* lz4: "last 5 bytes must be literals" */
    ORC_R_LZ4_ZERO_CAPACITY = 6,    /* lz4: Java returns -1 (zero-capacity output, input != {0}) */
    ORC_R_SNAPPY_TRUNCATED = 7,     /* snappy: "Input is truncated" */
    ORC_R_SNAPPY_VARINT_HIGHBIT = 8,/* snappy: "last byte of compressed length int has high bit set" */
    ORC_R_SNAPPY_NEG_LENGTH = 9,    /* snappy: "invalid compressed length" */
    ORC_R_SNAPPY_LEN_GT_CAP = 10,   /* snappy: IAE "Uncompressed length %s must be less than %s" */
    ORC_R_SNAPPY_LEN_MISMATCH = 11, /* snappy: "Recorded length is %s bytes but actual length ..." */
    ORC_R_MAX_OUTPUT_TOO_SMALL = 12,/* compressors: IAE "Max output length must be larger than" / "Output buffer must be at least" */
    ORC_R_MAX_INPUT_EXCEEDED = 13,  /* lz4: "Max input length exceeded" */
    /* zstd reasons start at 32, see zstd section */
    ORC_R_ZSTD_BASE = 32
};

#define ORC_STATUS(code, reason) (-(int64_t)((code) | ((reason) << 8)))

/* ---- LZ4 block: lz4/Lz4RawCompressor.java, lz4/Lz4RawDecompressor.java ---- */
int64_t orc_lz4_max_compressed_length(int64_t n);
int64_t orc_lz4_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap);
int64_t orc_lz4_decompress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, int64_t *err_offset);

/* ---- Snappy raw: snappy/SnappyRawCompressor.java, snappy/SnappyRawDecompressor.java ---- */
int64_t orc_snappy_max_compressed_length(int64_t n);
int64_t orc_snappy_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap);
int64_t orc_snappy_decompress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, int64_t *err_offset);
int64_t orc_snappy_uncompressed_length(const uint8_t *in, int64_t in_len, int64_t *err_offset);

/* ---- XXH64: zstd/XxHash64.java, xxhash/XxHash64JavaHasher.java ---- */
uint64_t orc_xxh64(const uint8_t *in, int64_t len, uint64_t seed);
uint64_t orc_xxh64_long(uint64_t value, uint64_t seed);
uint32_t orc_xxh32(const uint8_t *in, int64_t len, uint32_t seed);   /* xxhash/XxHash32JavaHasher.java:68-109 */

/* ---- Zstandard: zstd/ZstdFrameCompressor.java, zstd/ZstdFrameDecompressor.java (+ helpers) ---- */
int64_t orc_zstd_max_compressed_length(int64_t n);
int64_t orc_zstd_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap);
int64_t orc_zstd_decompress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, int64_t *err_offset);
int64_t orc_zstd_decompressed_size(const uint8_t *in, int64_t in_len, int64_t *err_offset);

/* ---- batch drivers used by tests and by bench.py's CPU baseline (OpenMP over independent blocks) ---- */
/* op: 0 lz4c 1 lz4d 2 snappyc 3 snappyd 4 zstdc 5 zstdd 6 xxh64 (out_len receives the hash) */
int64_t orc_batch(int32_t op, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                  uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap,
                  int64_t *out_len, int64_t n, int32_t threads);
/* same loop over the reference's bundled native libraries; fn = address of the native entry point for `op` (see batch_oracle.c) */
int64_t orc_native_batch(int32_t op, void *fn, const uint8_t *src_base, const int64_t *src_off, const int64_t *src_len,
                         uint8_t *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *out_len, int64_t n, int32_t threads);
int32_t orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
