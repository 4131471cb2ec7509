/*
 * xxh32_oracle.c -- CPU restatement of the reference's XXH32 (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows xxhash/XxHash32JavaHasher.java:68-109 (one-shot hash over a byte range), :343-366 (mix, the two
 * tail updates, finalShuffle).  The LZ4 frame format uses it for its header, block and content checksums
 * (lz4/Lz4FrameCompression.java:95,216,285,307).  Pinned by tests/test_oracle_golden.py against the
 * reference's known answers (T/xxhash/TestXxHash32.java:45-46) and its bundled libxxhash (oracle/_ref).
 */
#include "oracle.h"
#include <string.h>

#define Q1 0x9E3779B1u
#define Q2 0x85EBCA77u
#define Q3 0xC2B2AE3Du
#define Q4 0x27D4EB2Fu
#define Q5 0x165667B1u

static inline uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* XxHash32JavaHasher.java:343-346 mix */
static inline uint32_t mix(uint32_t cur, uint32_t value) { return rotl32(cur + value * Q2, 13) * Q1; }

/* XxHash32JavaHasher.java:68-109 hash(input, offset, length, seed) */
uint32_t orc_xxh32(const uint8_t *in, int64_t len, uint32_t seed)
{
    uint32_t hash;
    int64_t index = 0;
    if (len >= 16) {                                                   /* :76-90 */
        uint32_t v1 = seed + Q1 + Q2, v2 = seed + Q2, v3 = seed, v4 = seed - Q1;
        for (; index + 16 <= len; index += 16) {
            v1 = mix(v1, ld32(in + index));
            v2 = mix(v2, ld32(in + index + 4));
            v3 = mix(v3, ld32(in + index + 8));
            v4 = mix(v4, ld32(in + index + 12));
        }
        hash = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    }
    else {
        hash = seed + Q5;                                              /* :93 */
    }
    hash += (uint32_t) len;                                            /* :96 */
    for (; index + 4 <= len; index += 4) hash = rotl32(hash + ld32(in + index) * Q3, 17) * Q4;   /* :99-102, :348-351 */
    for (; index < len; index++) hash = rotl32(hash + in[index] * Q5, 11) * Q1;                  /* :104-107, :353-357 */
    hash ^= hash >> 15; hash *= Q2; hash ^= hash >> 13; hash *= Q3; hash ^= hash >> 16;          /* :359-366 finalShuffle */
    return hash;
}
