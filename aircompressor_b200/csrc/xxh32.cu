// xxh32.cu -- batched one-shot XXH32 for sm_100a (the checksum of the LZ4 frame format).
//
// Replaces XxHash32JavaHasher.hash(input, offset, length, seed) (xxhash/XxHash32JavaHasher.java:68-109, mix / tail updates /
// finalShuffle :343-366) and the native binding it stands in for, XXH32(input, length, seed) (xxhash/XxHash32Bindings.java).
// The LZ4 frame codec calls it for the header byte, per-block checksums and the content checksum
// (lz4/Lz4FrameCompression.java:95,216,285,307).  XXH32 has exactly four accumulator chains of 32-bit words, so one buffer is
// served by four lanes (lane `sub` owns accumulator v[sub]: word `sub` of every 16-byte stripe) and a warp hashes eight buffers
// at once; the tail (< 16 bytes) is one lane's work.  Buffers may start at any byte: a lane then assembles its word from the two
// aligned words around it (both contain bytes of the stripe, so nothing outside the aligned words of the buffer is read).
#include "acc_device.cuh"

namespace {

constexpr uint32_t Q1 = 0x9E3779B1u, Q2 = 0x85EBCA77u, Q3 = 0xC2B2AE3Du, Q4 = 0x27D4EB2Fu, Q5 = 0x165667B1u;
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int r) { return __funnelshift_l(v, v, r); }
__device__ __forceinline__ uint32_t mix32(uint32_t cur, uint32_t v) { return rotl32(cur + v * Q2, 13) * Q1; }

// Four consecutive lanes (gmask) hash one buffer; every lane of the group returns the hash.  Lanes without a buffer pass len = 0.
__device__ __forceinline__ uint32_t xxh32_group4(const uint8_t *in, int64_t len, uint32_t seed, int sub, unsigned gmask)
{
    uint32_t hash;
    const int64_t stripes = len >> 4;
    if (len >= 16) {
        uint32_t v = sub == 0 ? seed + Q1 + Q2 : sub == 1 ? seed + Q2 : sub == 2 ? seed : seed - Q1;
        const uint8_t *p = in + sub * 4;
        const uint32_t k = (uint32_t) ((uintptr_t) p & 3);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p - k);
        if (k == 0) {
            int64_t s = 0;
            for (; s + 4 <= stripes; s += 4) {
                const uint32_t a0 = q[(s + 0) * 4], a1 = q[(s + 1) * 4], a2 = q[(s + 2) * 4], a3 = q[(s + 3) * 4];
                v = mix32(v, a0); v = mix32(v, a1); v = mix32(v, a2); v = mix32(v, a3);
            }
            for (; s < stripes; s++) v = mix32(v, q[s * 4]);
        }
        else {
            for (int64_t s = 0; s < stripes; s++) v = mix32(v, __funnelshift_r(q[s * 4], q[s * 4 + 1], k * 8));
        }
        const int base = (lane_id() & ~3);
        const uint32_t v1 = __shfl_sync(gmask, v, base), v2 = __shfl_sync(gmask, v, base + 1), v3 = __shfl_sync(gmask, v, base + 2), v4 = __shfl_sync(gmask, v, base + 3);
        hash = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    }
    else hash = seed + Q5;
    hash += (uint32_t) len;
    if (sub == 0) {
        int64_t i = stripes << 4;
        for (; i + 4 <= len; i += 4) {
            const uint32_t w = (uint32_t) in[i] | ((uint32_t) in[i + 1] << 8) | ((uint32_t) in[i + 2] << 16) | ((uint32_t) in[i + 3] << 24);
            hash = rotl32(hash + w * Q3, 17) * Q4;
        }
        for (; i < len; i++) hash = rotl32(hash + in[i] * Q5, 11) * Q1;
        hash ^= hash >> 15; hash *= Q2; hash ^= hash >> 13; hash *= Q3; hash ^= hash >> 16;
    }
    return __shfl_sync(gmask, hash, lane_id() & ~3);
}

__global__ void __launch_bounds__(256) xxh32_kernel(AccBatch b, uint32_t seed)
{
    const int lane = lane_id();
    const int sub = lane & 3;           // accumulator index
    const int grp = lane >> 2;          // buffer slot inside the warp
    const int64_t warp_global = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps_total = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const unsigned gmask = 0xfu << (grp * 4);
    for (int64_t base = warp_global * 8; base < b.n; base += warps_total * 8) {
        const int64_t idx = base + grp;
        const bool active = idx < b.n;
        const uint8_t *in = active ? b.src + b.src_off[idx] : nullptr;
        const int64_t len = active ? b.src_len[idx] : 0;
        const uint32_t h = xxh32_group4(in, len, seed, sub, gmask);
        if (active && sub == 0) {
            b.out_len[idx] = (int64_t) h;     // zero-extended
            if (b.status) b.status[idx] = 0;
        }
    }
}

}  // namespace

void acc_launch_xxh32(const AccBatch &b, uint32_t seed, int sm_count, cudaStream_t st)
{
    int64_t warps = (b.n + 7) / 8;
    int64_t ctas = (warps + 7) / 8;
    int64_t max_ctas = (int64_t) sm_count * 8;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    xxh32_kernel<<<(unsigned) ctas, 256, 0, st>>>(b, seed);
}
