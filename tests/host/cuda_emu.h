// cuda_emu.h -- just enough of the CUDA device vocabulary to compile the record path of the LZ decoders
// (aircompressor_b200/csrc/lz_records.cuh and the codec headers) for the HOST, with OS threads as lanes.  Test
// infrastructure (tests/test_record_engine_emu.py): it checks the parse logic and the record hand-over on the CPU.
// A warp is 32 threads sharing one EmuWarp; __syncwarp() is a barrier over them.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <sched.h>

#define __device__
#define __global__
#define __host__
#define __noinline__ __attribute__((noinline))
#define __forceinline__ inline __attribute__((always_inline))
#define __constant__ const
#define __align__(n) __attribute__((aligned(n)))
#define __builtin_assume(x) ((void) 0)
#define __isGlobal(p) true

struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
typedef void *cudaStream_t;

struct __attribute__((aligned(8))) uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }

struct EmuWarp {
    pthread_barrier_t bar;
    int nlanes = 32;            // threads that play this warp (the barrier's count)
    volatile unsigned long long xchg[32];
};
extern thread_local EmuWarp *t_warp;
extern thread_local int t_lane;

static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&t_warp->bar); }
static inline int lane_id_emu() { return t_lane; }

// warp collectives over all 32 lanes (the engine only uses full masks): exchange through the warp's array
static inline unsigned long long emu_xchg(unsigned long long v, int src)
{
    t_warp->xchg[t_lane] = v;
    __syncwarp();
    const unsigned long long r = t_warp->xchg[src & 31];
    __syncwarp();
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return (T) emu_xchg((unsigned long long) (long long) v, src); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta) { const int src = t_lane - (int) delta; return (T) emu_xchg((unsigned long long) v, src < 0 ? t_lane : src); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) { const int src = t_lane + (int) delta; return (T) emu_xchg((unsigned long long) v, src > 31 ? t_lane : src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return (T) emu_xchg((unsigned long long) v, t_lane ^ m); }
static inline unsigned __ballot_sync(unsigned, bool p)
{
    t_warp->xchg[t_lane] = p ? 1 : 0;
    __syncwarp();
    unsigned m = 0;
    for (int i = 0; i < t_warp->nlanes; i++) m |= (unsigned) (t_warp->xchg[i] & 1) << i;
    __syncwarp();
    return m;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    t_warp->xchg[t_lane] = v;
    __syncwarp();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned) t_warp->xchg[i];
    __syncwarp();
    return m;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0; }
static inline bool __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, !p) == 0; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int) v); }
[[noreturn]] static inline void emu_unsupported(const char *what) { fprintf(stderr, "cuda_emu: %s reached\n", what); abort(); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *(const volatile T *) p; }
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t) (((((uint64_t) hi) << 32) | lo) >> (sh > 32 ? 32 : sh)); }
static inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t) ((((((uint64_t) hi) << 32) | lo) << (sh > 32 ? 32 : sh)) >> 32); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t) y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t) ((v >> (((s >> (4 * i)) & 7) * 8)) & 0xFF) << (8 * i);
    return r;
}
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t) (((((uint64_t) hi) << 32) | lo) >> (sh & 31)); }
