// xxh32.cu -- batched one-shot XXH32 for sm_100a (the checksum of the LZ4 frame format).
//
// Replaces XxHash32JavaHasher.hash(input, offset, length, seed) (xxhash/XxHash32JavaHasher.java:68-109, mix / tail updates /
// finalShuffle :343-366) and the native binding it stands in for, XXH32(input, length, seed) (xxhash/XxHash32Bindings.java).
// The LZ4 frame codec calls it for the header byte, per-block checksums and the content checksum
// (lz4/Lz4FrameCompression.java:95,216,285,307).  XXH32 has exactly four accumulator chains of 32-bit words, so one buffer is
// served by four lanes (lane `sub` owns accumulator v[sub]: word `sub` of every 16-byte stripe) and a warp hashes eight buffers
// at once; the tail (< 16 bytes) is one lane's work.  Buffers may start at any byte.
#include "acc_device.cuh"

namespace {

constexpr uint32_t Q1 = 0x9E3779B1u, Q2 = 0x85EBCA77u, Q3 = 0xC2B2AE3Du, Q4 = 0x27D4EB2Fu, Q5 = 0x165667B1u;
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int r) { return __funnelshift_l(v, v, r); }
__device__ __forceinline__ uint32_t mix32(uint32_t cur, uint32_t v) { return rotl32(cur + v * Q2, 13) * Q1; }

// The batch kernel stages the input in shared memory: a warp serves eight buffers at a time, and for each of them all 32
// lanes fetch one 512-byte chunk with 16-byte loads (one fully used wavefront per 512 bytes instead of one per 16), then the
// buffer's four lanes read their words back from shared memory (rows 544 bytes apart: the eight groups hit disjoint banks).
// Any alignment: the chunk starts at the 16-byte aligned address below the buffer, a lane assembles its word from two shared
// words when the buffer is not 4-byte aligned.  The accumulator chains and the tail are those of xxh32_group4.
constexpr int kChunk = 512, kRow = kChunk + 32, kWarps = 8;

__global__ void __launch_bounds__(kWarps * 32) xxh32_kernel(AccBatch b, uint32_t seed)
{
    __shared__ __align__(16) uint8_t stage[kWarps][8][kRow];
    const int lane = lane_id();
    const int sub = lane & 3;           // accumulator index
    const int grp = lane >> 2;          // buffer slot inside the warp
    const int warp = threadIdx.x >> 5;
    const int64_t warp_global = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps_total = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const unsigned gmask = 0xfu << (grp * 4);
    for (int64_t base = warp_global * 8; base < b.n; base += warps_total * 8) {
        const int64_t idx = base + grp;
        const bool active = idx < b.n;
        const uint8_t *in = active ? b.src + b.src_off[idx] : nullptr;
        const int64_t len = active ? b.src_len[idx] : 0;
        const int64_t stripes = len >> 4;
        const uint32_t k16 = (uint32_t) ((uintptr_t) in & 15);
        const uint8_t *abase = in - k16;                                  // 16-byte aligned
        const int64_t need = stripes ? k16 + (stripes << 4) : 0;          // bytes from abase that hold stripes
        uint32_t v = sub == 0 ? seed + Q1 + Q2 : sub == 1 ? seed + Q2 : sub == 2 ? seed : seed - Q1;
        const int64_t my_chunks = (stripes + 31) >> 5;                    // 32 stripes per chunk
        int64_t max_chunks = my_chunks;
        for (int o = 16; o; o >>= 1) { const int64_t t = __shfl_xor_sync(kFull, max_chunks, o); max_chunks = t > max_chunks ? t : max_chunks; }
        for (int64_t c = 0; c < max_chunks; c++) {
            // ---- stage chunk c of the eight buffers: bytes [512 c, 512 c + 544) from abase, as far as they hold stripes
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const uint8_t *ab = reinterpret_cast<const uint8_t *>(__shfl_sync(kFull, reinterpret_cast<uintptr_t>(abase), g * 4));
                const int64_t nd = __shfl_sync(kFull, need, g * 4);
                const int64_t o0 = c * kChunk + lane * 16;
                if (o0 < nd) *reinterpret_cast<uint4 *>(&stage[warp][g][lane * 16]) = *reinterpret_cast<const uint4 *>(ab + o0);
                const int64_t o1 = c * kChunk + kChunk + lane * 16;       // the 32 bytes behind the chunk: a misaligned last stripe ends there
                if (lane < 2 && o1 < nd) *reinterpret_cast<uint4 *>(&stage[warp][g][kChunk + lane * 16]) = *reinterpret_cast<const uint4 *>(ab + o1);
            }
            __syncwarp();
            // ---- my group's stripes of this chunk
            if (c < my_chunks) {
                const int n = (int) (stripes - (c << 5) < 32 ? stripes - (c << 5) : 32);
                const uint8_t *row = &stage[warp][grp][0];
                const uint32_t off = k16 + (uint32_t) sub * 4, k = off & 3;
                const uint32_t *w = reinterpret_cast<const uint32_t *>(row + (off - k));
                if (k == 0) {
                    int j = 0;
                    for (; j + 4 <= n; j += 4) {
                        const uint32_t a0 = w[j * 4], a1 = w[j * 4 + 4], a2 = w[j * 4 + 8], a3 = w[j * 4 + 12];
                        v = mix32(v, a0); v = mix32(v, a1); v = mix32(v, a2); v = mix32(v, a3);
                    }
                    for (; j < n; j++) v = mix32(v, w[j * 4]);
                }
                else {
                    for (int j = 0; j < n; j++) v = mix32(v, __funnelshift_r(w[j * 4], w[j * 4 + 1], k * 8));
                }
            }
            __syncwarp();
        }
        uint32_t hash;
        if (len >= 16) {
            const int g0 = lane & ~3;
            const uint32_t v1 = __shfl_sync(gmask, v, g0), v2 = __shfl_sync(gmask, v, g0 + 1), v3 = __shfl_sync(gmask, v, g0 + 2), v4 = __shfl_sync(gmask, v, g0 + 3);
            hash = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
        }
        else hash = seed + Q5;
        hash += (uint32_t) len;
        if (active && sub == 0) {
            int64_t i = stripes << 4;
            for (; i + 4 <= len; i += 4) {
                const uint32_t x = (uint32_t) in[i] | ((uint32_t) in[i + 1] << 8) | ((uint32_t) in[i + 2] << 16) | ((uint32_t) in[i + 3] << 24);
                hash = rotl32(hash + x * Q3, 17) * Q4;
            }
            for (; i < len; i++) hash = rotl32(hash + in[i] * Q5, 11) * Q1;
            hash ^= hash >> 15; hash *= Q2; hash ^= hash >> 13; hash *= Q3; hash ^= hash >> 16;
            b.out_len[idx] = (int64_t) hash;     // zero-extended
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
