/*
 * xxh64_oracle.c -- CPU restatement of the reference's XXH64 (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows  zstd/XxHash64.java:182-290  (one-shot hash used for the zstd frame checksum) and
 * xxhash/XxHash64JavaHasher.java:66-124 (public one-shot hash, hash(long value, long seed)).
 */
#include "oracle.h"
#include <string.h>

#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* XxHash64.java:250-253 mix */
static inline uint64_t mix(uint64_t cur, uint64_t value) { return rotl(cur + value * P2, 31) * P1; }
/* XxHash64.java:255-259 update (merge of one lane accumulator) */
static inline uint64_t merge(uint64_t hash, uint64_t value) { return (hash ^ mix(0, value)) * P1 + P4; }
/* XxHash64.java:280-289 finalShuffle */
static inline uint64_t avalanche(uint64_t h)
{
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* XxHash64.java:182-249 hash(seed, base, address, length) */
uint64_t orc_xxh64(const uint8_t *in, int64_t len, uint64_t seed)
{
    uint64_t hash;
    int64_t index = 0;
    if (len >= 32) {                                                   /* updateBody :222-248 */
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        for (; index + 32 <= len; index += 32) {
            v1 = mix(v1, ld64(in + index));
            v2 = mix(v2, ld64(in + index + 8));
            v3 = mix(v3, ld64(in + index + 16));
            v4 = mix(v4, ld64(in + index + 24));
        }
        hash = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        hash = merge(hash, v1); hash = merge(hash, v2); hash = merge(hash, v3); hash = merge(hash, v4);
    }
    else {
        hash = seed + P5;
    }
    hash += (uint64_t) len;
    /* updateTail :201-220, :261-278 */
    for (; index <= len - 8; index += 8) hash = rotl(hash ^ mix(0, ld64(in + index)), 27) * P1 + P4;
    if (index <= len - 4) { hash = rotl(hash ^ ((uint64_t) ld32(in + index) * P1), 23) * P2 + P3; index += 4; }
    for (; index < len; index++) hash = rotl(hash ^ ((uint64_t) in[index] * P5), 11) * P1;
    return avalanche(hash);
}

/* XxHash64JavaHasher.java:66-71 hash(long value, long seed) */
uint64_t orc_xxh64_long(uint64_t value, uint64_t seed)
{
    uint64_t hash = seed + P5 + 8;
    hash = rotl(hash ^ mix(0, value), 27) * P1 + P4;
    return avalanche(hash);
}
