// cuda_emu.h -- just enough of the CUDA device vocabulary to compile the streaming decode engine
// (aircompressor_b200/csrc/lz_stream.cuh and the codec headers) for the HOST, with OS threads as lanes.  Test
// infrastructure (tests/test_stream_engine_emu.py): it checks the queue / ring protocol and the parse logic on the CPU.
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
#define __forceinline__ inline __attribute__((always_inline))
#define __constant__ const
#define __align__(n) __attribute__((aligned(n)))
#define __builtin_assume(x) ((void) 0)
#define __isGlobal(p) true

struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
typedef void *cudaStream_t;

struct EmuWarp {
    pthread_barrier_t bar;
    volatile uint32_t bcast;
};
extern thread_local EmuWarp *t_warp;
extern thread_local int t_lane;

static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&t_warp->bar); }
static inline int lane_id_emu() { return t_lane; }

// warp shuffles / votes are only reached by the round-1 step decoders, which the emulation never selects
[[noreturn]] static inline void emu_unsupported(const char *what) { fprintf(stderr, "cuda_emu: %s reached\n", what); abort(); }
static inline uint32_t __shfl_sync(unsigned, uint32_t, int) { emu_unsupported("__shfl_sync"); }
static inline bool __any_sync(unsigned, bool) { emu_unsupported("__any_sync"); }
static inline unsigned __ballot_sync(unsigned, bool) { emu_unsupported("__ballot_sync"); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *(const volatile T *) p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t) (((((uint64_t) hi) << 32) | lo) >> (sh & 31)); }
