// acc_device.cuh -- shared device-side definitions for libaircompress_cuda.so (sm_100a).
//
// Data layout in HBM for every batch kernel: one packed source buffer and one packed destination
// buffer, plus four int64 index arrays (src_off, src_len, dst_off, dst_cap) and two result arrays
// (out_len int64, status int32).  Block i is independent of every other block -- the reference's unit
// of work is one Compressor/Decompressor call (Compressor.java:18-36, Decompressor.java:18-31).
#pragma once
#include <cstdint>
#ifdef LZS_EMU
#include "cuda_emu.h"   // host emulation of the record path (tests/host)
#else
#include <cuda_runtime.h>
#endif
#include "../../include/aircompress_cuda.h"

struct AccBatch {
    const uint8_t *src;
    const int64_t *src_off;
    const int64_t *src_len;
    uint8_t *dst;
    const int64_t *dst_off;
    const int64_t *dst_cap;
    int64_t *out_len;
    int32_t *status;
    int64_t n;
    unsigned int *work_counter;  // zeroed before launch; persistent CTAs claim block indices from it
};

#define ACC_STATUS(code, reason) ((int32_t) ((code) | ((reason) << 8)))

static constexpr int kWarp = 32;
static constexpr unsigned kFull = 0xffffffffu;

#ifdef LZS_EMU
__device__ __forceinline__ int lane_id() { return lane_id_emu(); }
#else
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

__device__ __forceinline__ uint32_t ld_u16le(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8); }

// Unaligned little-endian loads assembled from aligned 32-bit words (global or shared).
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p)
{
    uintptr_t a = (uintptr_t) p;
    uint32_t sh = (uint32_t) (a & 3);
    const uint32_t *w = (const uint32_t *) (a - sh);
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    uint32_t hi = w[1];
    return __funnelshift_r(lo, hi, sh * 8);
}

__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t *p)
{
    uintptr_t a = (uintptr_t) p;
    uint32_t sh = (uint32_t) (a & 3);
    const uint32_t *w = (const uint32_t *) (a - sh);
    uint32_t w0 = w[0], w1 = w[1];
    uint32_t w2 = sh ? w[2] : 0;
    uint32_t lo = __funnelshift_r(w0, w1, sh * 8);
    uint32_t hi = __funnelshift_r(w1, w2, sh * 8);
    return (uint64_t) lo | ((uint64_t) hi << 32);
}

// Warp-cooperative copy of n bytes, src and dst do not overlap (or src is entirely final data).
// Short runs (the common case on LZ sequences) take one predicated byte move per lane; long runs use
// 16-byte stores with the source re-aligned through funnel shifts.
// NOTE: no __restrict__ on src -- match copies read bytes this kernel wrote earlier, which must not go
// through the non-coherent (ld.global.nc) path.  kCg = true reads through L2 only (ld.global.cg), for
// sources that were produced with atomics (which bypass L1).
template <bool kCg = false>
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, int64_t n, int lane)
{
    auto ld8 = [](const uint8_t *p) -> uint8_t { return kCg ? __ldcg(p) : *p; };
    auto ld32 = [](const uint32_t *p) -> uint32_t { return kCg ? __ldcg(p) : *p; };
    if (n <= 64) {
        if (lane < n) dst[lane] = ld8(src + lane);
        if (lane + 32 < n) dst[lane + 32] = ld8(src + lane + 32);
        return;
    }
    int head = (int) ((16 - ((uintptr_t) dst & 15)) & 15);
    if (lane < head) dst[lane] = ld8(src + lane);
    dst += head; src += head; n -= head;
    int64_t nvec = n >> 4;
    uintptr_t sa = (uintptr_t) src;
    uint32_t sh = (uint32_t) (sa & 3);
    const uint32_t *s32 = (const uint32_t *) (sa - sh);
    uint4 *d16 = (uint4 *) dst;
    for (int64_t v = lane; v < nvec; v += 32) {
        const uint32_t *s = s32 + v * 4;
        uint32_t w0 = ld32(s), w1 = ld32(s + 1), w2 = ld32(s + 2), w3 = ld32(s + 3);
        uint32_t w4 = sh ? ld32(s + 4) : 0;
        uint4 o;
        o.x = __funnelshift_r(w0, w1, sh * 8);
        o.y = __funnelshift_r(w1, w2, sh * 8);
        o.z = __funnelshift_r(w2, w3, sh * 8);
        o.w = __funnelshift_r(w3, w4, sh * 8);
        d16[v] = o;
    }
    int64_t done = nvec << 4;
    int tail = (int) (n - done);
    if (lane < tail) dst[done + lane] = ld8(src + done + lane);
}

// floor(65536 / d) + 1: (m * kRcp16[d]) >> 16 == m / d for m < 32
static __constant__ uint32_t kRcp16[32] = {0, 65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554, 5958, 5462, 5042, 4682, 4370,
                                    4097, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115};

// Warp-cooperative LZ77 match copy inside the output buffer with forward byte-copy semantics:
// dst[i] = dst[i - offset] for i in [0, len).  Callers __syncwarp() first so that earlier stores of
// other lanes are visible.  offset >= 1.
__device__ __forceinline__ void warp_match_copy(uint8_t *dst, int64_t offset, int64_t len, int lane)
{
    const uint8_t *src = dst - offset;
    if (offset >= len) {
        warp_copy(dst, src, len, lane);
    }
    else if (offset >= 32) {
        // every 32-byte step only reads bytes produced by earlier steps
        for (int64_t base = 0; base < len; base += 32) {
            int64_t i = base + lane;
            if (i < len) dst[i] = src[i];
            __syncwarp();
        }
    }
    else {
        // periodic pattern: all reads come from the `offset` bytes that precede dst
        int off = (int) offset;
        int m = lane % off;
        int step = 32 % off;
        for (int64_t i = lane; i < len; i += 32) {
            dst[i] = src[m];
            m += step;
            if (m >= off) m -= off;
        }
    }
}

// kernel launchers implemented in the per-codec translation units
void acc_launch_lz4_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st, void *scratch, unsigned int *second_counter);
int64_t acc_lz_records_scratch_bytes(int64_t n);   // record path of the LZ4 / Snappy decoders (lz_records.cuh); scratch == nullptr: step decoder alone
void acc_launch_lz4_compress(const AccBatch &b, int sm_count, cudaStream_t st, unsigned int *second_counter);
void acc_launch_snappy_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st, void *scratch, unsigned int *second_counter);
void acc_launch_snappy_compress(const AccBatch &b, int sm_count, cudaStream_t st);
void acc_launch_xxh64(const AccBatch &b, uint64_t seed, int sm_count, cudaStream_t st);
void acc_launch_xxh32(const AccBatch &b, uint32_t seed, int sm_count, cudaStream_t st);
void acc_launch_zstd_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st, void *scratch, int64_t scratch_bytes);
void acc_launch_zstd_compress(const AccBatch &b, int sm_count, cudaStream_t st, void *scratch, int64_t scratch_bytes);
