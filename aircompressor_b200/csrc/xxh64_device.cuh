// xxh64_device.cuh -- XXH64 device routines shared by the batch hash kernel and the zstd frame
// checksum (zstd/ZstdFrameCompressor.java:123-134, zstd/ZstdFrameDecompressor.java:194-206).
// Arithmetic follows zstd/XxHash64.java:201-289.
#pragma once
#include <cstdint>
#include "acc_device.cuh"

namespace xxh {
constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t P2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t P3 = 0x165667B19E3779F9ULL;
constexpr uint64_t P4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t P5 = 0x27D4EB2F165667C5ULL;
__device__ __forceinline__ uint64_t rotl(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
__device__ __forceinline__ uint64_t mix(uint64_t cur, uint64_t v) { return rotl(cur + v * P2, 31) * P1; }
__device__ __forceinline__ uint64_t merge(uint64_t h, uint64_t v) { return (h ^ mix(0, v)) * P1 + P4; }
__device__ __forceinline__ uint64_t avalanche(uint64_t h)
{
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
}  // namespace xxh

// Four consecutive lanes (gmask) hash one buffer; lane `sub` (0..3) owns accumulator v[sub].
// Every lane of the group returns the final hash.  Lanes whose buffer is empty/absent pass len = 0.
__device__ __forceinline__ uint64_t xxh64_group4(const uint8_t *in, int64_t len, uint64_t seed, int sub, unsigned gmask)
{
    using namespace xxh;
    uint64_t hash;
    const int64_t stripes = len >> 5;
    if (len >= 32) {
        uint64_t v = sub == 0 ? seed + P1 + P2 : sub == 1 ? seed + P2 : sub == 2 ? seed : seed - P1;
        const uint8_t *p = in + sub * 8;
        if ((((uintptr_t) in) & 7) == 0) {
            const uint64_t *q = (const uint64_t *) p;
            int64_t s = 0;
            for (; s + 4 <= stripes; s += 4) {
                uint64_t a0 = q[(s + 0) * 4], a1 = q[(s + 1) * 4], a2 = q[(s + 2) * 4], a3 = q[(s + 3) * 4];
                v = mix(v, a0); v = mix(v, a1); v = mix(v, a2); v = mix(v, a3);
            }
            for (; s < stripes; s++) v = mix(v, q[s * 4]);
        }
        else {
            // any other alignment (blocks packed behind a ragged one): aligned 8-byte words, each value assembled from two of
            // them -- the same number of independent loads in flight as the aligned loop instead of a chain of 4-byte loads
            const uintptr_t a = (uintptr_t) p;
            const uint32_t sh = (uint32_t) (a & 7) * 8;                 // != 0 here
            const uint64_t *q = (const uint64_t *) (a & ~(uintptr_t) 7);
            int64_t s = 0;
            for (; s + 4 <= stripes; s += 4) {
                const uint64_t l0 = q[(s + 0) * 4], h0 = q[(s + 0) * 4 + 1], l1 = q[(s + 1) * 4], h1 = q[(s + 1) * 4 + 1];
                const uint64_t l2 = q[(s + 2) * 4], h2 = q[(s + 2) * 4 + 1], l3 = q[(s + 3) * 4], h3 = q[(s + 3) * 4 + 1];
                v = mix(v, (l0 >> sh) | (h0 << (64 - sh))); v = mix(v, (l1 >> sh) | (h1 << (64 - sh)));
                v = mix(v, (l2 >> sh) | (h2 << (64 - sh))); v = mix(v, (l3 >> sh) | (h3 << (64 - sh)));
            }
            for (; s < stripes; s++) v = mix(v, (q[s * 4] >> sh) | (q[s * 4 + 1] << (64 - sh)));
        }
        // gather the four accumulators of the group
        const int lane = lane_id();
        const int g0 = lane & ~3;
        uint64_t v1 = __shfl_sync(gmask, v, g0 + 0);
        uint64_t v2 = __shfl_sync(gmask, v, g0 + 1);
        uint64_t v3 = __shfl_sync(gmask, v, g0 + 2);
        uint64_t v4 = __shfl_sync(gmask, v, g0 + 3);
        hash = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        hash = merge(hash, v1); hash = merge(hash, v2); hash = merge(hash, v3); hash = merge(hash, v4);
    }
    else {
        hash = seed + P5;
    }
    hash += (uint64_t) len;
    int64_t index = stripes << 5;
    for (; index <= len - 8; index += 8) hash = rotl(hash ^ mix(0, ld_u64_unaligned(in + index)), 27) * P1 + P4;
    if (index <= len - 4) { hash = rotl(hash ^ ((uint64_t) ld_u32_unaligned(in + index) * P1), 23) * P2 + P3; index += 4; }
    for (; index < len; index++) hash = rotl(hash ^ ((uint64_t) in[index] * P5), 11) * P1;
    return avalanche(hash);
}
