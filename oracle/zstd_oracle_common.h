/*
 * zstd_oracle_common.h -- constants and helpers shared by the Zstandard oracle files
 * (TEST INFRASTRUCTURE, see oracle.h).  Values restate zstd/Constants.java:23-78,
 * zstd/ZstdFrameDecompressor.java:68-83, zstd/SequenceEncoder.java:36-56 and
 * zstd/FseCompressionTable.java:133-154.
 */
#ifndef ZSTD_ORACLE_COMMON_H
#define ZSTD_ORACLE_COMMON_H
#include <stdint.h>
#include <string.h>

#define ZO_MAX_BLOCK (128 * 1024)

/* reasons (status >> 8) for the zstd path; texts are the reference's exception messages */
enum {
    ZR_NOT_ENOUGH_INPUT = 32,        /* "Not enough input bytes" */
    ZR_INVALID_BLOCK_TYPE = 33,      /* "Invalid block type" */
    ZR_OUTPUT_TOO_SMALL = 34,        /* "Output buffer too small" (MalformedInputException in the Java decoder) */
    ZR_CORRUPTED = 35,               /* "Input is corrupted" */
    ZR_BAD_MAGIC = 36,               /* "Invalid magic prefix: ..." */
    ZR_V07_FORMAT = 37,              /* "Data encoded in unsupported ZSTD v0.7 format" */
    ZR_DICTIONARY = 38,              /* "Custom dictionaries not supported" */
    ZR_WINDOW_TOO_LARGE = 39,        /* "Window size too large (not yet supported)" */
    ZR_BLOCK_TOO_SMALL = 40,         /* "Compressed block size too small" */
    ZR_EXPECTED_TABLE = 41,          /* "Expected match length table to be present" */
    ZR_DICTIONARY_CORRUPTED = 42,    /* "Dictionary is corrupted" (treeless literals without a table) */
    ZR_BLOCK_EXCEEDS_MAX = 44,       /* "Block exceeds maximum size" */
    ZR_OUTPUT_EXCEEDS_BLOCK = 45,    /* "Output exceeds maximum block size" */
    ZR_VALUE_EXCEEDS_MAX = 46,       /* "Value exceeds expected maximum value" */
    ZR_BITSTREAM_EMPTY = 47,         /* "Bitstream is empty" */
    ZR_BITSTREAM_NO_END_MARK = 48,   /* "Bitstream end mark not present" */
    ZR_NOT_ALL_SEQUENCES = 49,       /* "Not all sequences were consumed" */
    ZR_BITSTREAM_NOT_CONSUMED = 50,  /* "Bit stream is not fully consumed" */
    ZR_FSE_TABLE_TOO_LARGE = 51,     /* "FSE table size exceeds maximum allowed size" */
    ZR_SYMBOL_TOO_LARGE = 52,        /* "Symbol larger than max value" */
    ZR_TOO_MANY_SYMBOLS = 53,        /* "Max symbol value too large (too many symbols for FSE)" */
    ZR_BAD_CHECKSUM = 54,            /* "Bad checksum. Expected: ..., actual: ..." */
    ZR_FSE_OUTPUT_TOO_SMALL = 55     /* "Output buffer is too small" (Huffman weights) */
};

static inline uint64_t zo_ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t zo_ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t zo_ld16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline void zo_st16(uint8_t *p, uint32_t v) { uint16_t s = (uint16_t) v; memcpy(p, &s, 2); }
static inline void zo_st32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static inline void zo_st64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }
static inline void zo_st24(uint8_t *p, uint32_t v) { p[0] = (uint8_t) v; p[1] = (uint8_t) (v >> 8); p[2] = (uint8_t) (v >> 16); }

/* Util.highestBit :27-30 -- 31 - numberOfLeadingZeros(value); Java gives -1 for 0 */
static inline int zo_highbit(uint32_t v) { return v ? 31 - __builtin_clz(v) : -1; }

/* forward LZ77 copy with byte-copy semantics */
static inline void zo_match_copy(uint8_t *dst, const uint8_t *src, int64_t len)
{
    if (dst - src >= 8) {
        while (len >= 8) { uint64_t v; memcpy(&v, src, 8); memcpy(dst, &v, 8); dst += 8; src += 8; len -= 8; }
    }
    while (len-- > 0) *dst++ = *src++;
}

/* FseCompressionTable.spreadSymbols :138-154 */
static inline int zo_spread_symbols(const int16_t *norm, int max_symbol, int table_size, int high_threshold, uint8_t *symbols)
{
    int mask = table_size - 1;
    int step = (table_size >> 1) + (table_size >> 3) + 3;
    int position = 0;
    for (int s = 0; s <= max_symbol; s++) {
        for (int i = 0; i < norm[s]; i++) {
            symbols[position] = (uint8_t) s;
            do { position = (position + step) & mask; } while (position > high_threshold);
        }
    }
    return position;
}

/* Constants.java:66-78 */
static const uint8_t ZO_LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint8_t ZO_ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                       1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
/* ZstdFrameDecompressor.java:79-83: (1 << code) - 3 for code >= 2 */
static const int32_t ZO_OF_BASE[32] = {0, 1, 1, 5, 0xD, 0x1D, 0x3D, 0x7D, 0xFD, 0x1FD, 0x3FD, 0x7FD, 0xFFD, 0x1FFD, 0x3FFD, 0x7FFD, 0xFFFD,
                                       0x1FFFD, 0x3FFFD, 0x7FFFD, 0xFFFFD, 0x1FFFFD, 0x3FFFFD, 0x7FFFFD, 0xFFFFFD, 0x1FFFFFD, 0x3FFFFFD,
                                       0x7FFFFFD, 0xFFFFFFD, 0, 0, 0};
/* SequenceEncoder.java:36-56 predefined distributions (RFC 8878 section 3.1.1.3.2.2) */
static const int16_t ZO_DEFAULT_LL_NORM[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ZO_DEFAULT_ML_NORM[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                               1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t ZO_DEFAULT_OF_NORM[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
#endif
