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
            // any other alignment (blocks packed behind a ragged one): every lane loads ONE aligned 8-byte word per stripe
            // (lane `sub` the word that holds the low part of its value); the high part is the word of the next lane of the
            // group -- for the last lane the first word of the NEXT stripe, which lane 0 has already loaded one iteration
            // ahead.  Same global traffic and the same number of loads as the aligned loop, two shuffles per value on top.
            const uintptr_t a = (uintptr_t) p;
            const uint32_t sh = (uint32_t) (a & 7) * 8;                 // != 0 here
            const uint64_t *q = (const uint64_t *) (a & ~(uintptr_t) 7);
            const int lane = lane_id();
            const int nb = (lane & ~3) | ((sub + 1) & 3);                // the neighbour that holds my high part
            uint64_t w0 = q[0];
            int64_t s = 0;
            for (; s + 4 <= stripes; s += 4) {                          // four loads in flight per lane, like the aligned loop
                const uint64_t w1 = q[(s + 1) * 4], w2 = q[(s + 2) * 4], w3 = q[(s + 3) * 4];
                // behind the last stripe only lane 0's word is needed (it holds the last byte of lane 3's value, so it is inside
                // the buffer's aligned extent); the other lanes do not read there
                const uint64_t w4 = (s + 4 < stripes || sub == 0) ? q[(s + 4) * 4] : 0;
                const uint64_t c0 = __shfl_sync(gmask, w0, nb), c1 = __shfl_sync(gmask, w1, nb), c2 = __shfl_sync(gmask, w2, nb),
                               c3 = __shfl_sync(gmask, w3, nb), c4 = __shfl_sync(gmask, w4, nb);
                const uint64_t h0 = sub == 3 ? c1 : c0, h1 = sub == 3 ? c2 : c1, h2 = sub == 3 ? c3 : c2, h3 = sub == 3 ? c4 : c3;
                v = mix(v, (w0 >> sh) | (h0 << (64 - sh))); v = mix(v, (w1 >> sh) | (h1 << (64 - sh)));
                v = mix(v, (w2 >> sh) | (h2 << (64 - sh))); v = mix(v, (w3 >> sh) | (h3 << (64 - sh)));
                w0 = w4;
            }
            for (; s < stripes; s++) {
                const uint64_t nxt = (s + 1 < stripes || sub == 0) ? q[(s + 1) * 4] : 0;
                const uint64_t from_cur = __shfl_sync(gmask, w0, nb), from_nxt = __shfl_sync(gmask, nxt, nb);
                v = mix(v, (w0 >> sh) | ((sub == 3 ? from_nxt : from_cur) << (64 - sh)));
                w0 = nxt;
            }
        }
        // gather the four accumulators of the group
        const int g0 = lane_id() & ~3;
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
