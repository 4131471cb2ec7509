// lz_stream.cuh -- the two-phase LZ77 decode engine shared by the LZ4 and Snappy block decoders (sm_100a).
//
// A byte-oriented LZ decoder has two very different halves: walking the token grammar is a serial chain of tiny
// dependent steps (one token decides where the next one starts), while producing the bytes is wide copy work.  A warp
// that does both for ONE block spends most of its instructions on the serial half with one useful lane.  Here every warp
// owns up to 32 blocks and alternates between two phases:
//
//   PARSE   every lane walks the grammar of ITS OWN block for up to kRecPerLane sequences: a warp instruction advances up
//           to 32 independent token chains.  A lane reads its compressed bytes through a private 32-byte shared-memory
//           window (refilled with two aligned 16-byte loads: one global wavefront per ~4 sequences instead of one per
//           byte) and writes 8-byte RECORDS {literal length, match length, offset, header bytes skipped} to its row of
//           the warp's shared-memory record table;
//   EXECUTE the warp takes the rows one block at a time and produces the bytes: a whole short sequence is ONE load (a
//           literal of the input or an older output byte, selected per lane) and ONE store per lane; long literal runs and
//           long / overlapping matches use the 16-byte warp copies of acc_device.cuh.
//
// The parse lanes only accept what they can prove the reference decoder takes on its normal path
// (Lz4RawDecompressor.java:59-195, SnappyRawDecompressor.java:70-220); everything else -- the end-of-block rules,
// malformed input, very long lengths -- ends the fast path at a token boundary, and the warp finishes that block from
// the (input, output) position of the token with the exact restatement of the Java loop (lz4_decode_v1.cuh general path /
// snappy_decode_from).  Accept/reject decisions, error offsets and output bytes are therefore those of the general path
// by construction.
//
// (A first version of this engine gave the two halves to different warps -- one parse warp feeding 31 execute warps
// through shared-memory queues, input staged by cp.async.bulk.  It was bit-exact but 4x slower than round 1: a single
// warp issues a dependent instruction every ~5 cycles, so one parse warp capped the SM at 0.25 B/cycle however many
// lanes it used.  profiles/README.md keeps the numbers.)
//
// The same source compiles for the host with LZS_EMU defined (tests/host/lzs_emu.cpp: OS threads as lanes), which is
// how the parse logic and the phase hand-over are checked on the CPU (tests/test_stream_engine_emu.py).
#pragma once
#include "acc_device.cuh"

namespace lzs {

constexpr int kRecPerLane = 16;                 // records a lane may produce per parse phase
constexpr int kRecStride = kRecPerLane + 1;     // row pitch in records (odd: rows start in different banks)
constexpr int kWinBytes = 32;                   // per-lane input window
constexpr int kWinStride = 48;                  // window pitch (16-byte aligned, spreads the lanes over the banks)
constexpr uint32_t kNoOffset = 0xffffffu;       // offset field of a literal-only record
constexpr int kMaxLitPiece = 4095;              // ll field: 12 bits (longer runs are cut into pieces)
constexpr uint32_t kMaxMatch = (1u << 20) - 1;  // ml field: 20 bits
constexpr uint32_t kMaxOffset = (1u << 24) - 2; // off field: 24 bits
constexpr int kMaxSkip = 255;                   // skip field: 8 bits
constexpr uint32_t kWholeBlock = 0xffffffffu;   // restart position meaning "the general path decodes the whole block"

struct __align__(16) WarpSmem {
    uint8_t win[32 * kWinStride];
    uint2 rec[32 * kRecStride];
};

#ifndef LZS_EMU
__device__ __forceinline__ uint32_t claim_block(unsigned int *counter) { return atomicAdd(counter, 1u); }
#else
inline uint32_t claim_block(unsigned int *counter) { return __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED); }
#endif

enum ParseResult { kBudgetUsed = 0, kFallback = 2 };

// What a codec's parse loop sees of its lane's block.  All positions are relative to `in`.
struct ParseCtx {
    const uint8_t *in;        // first byte the grammar walk reads (behind the codec's preamble, if it has one)
    int32_t in_len, out_cap;
    uint8_t *win;             // this lane's 32-byte window
    uint2 *rec;               // this lane's record row
    uint32_t win_tag;         // the aligned 32-byte chunk the window holds; ~0: none
    uint32_t head;            // (address of in) & 31
    int32_t prev_lit_end;     // input position behind the literals of the previous record
    int n_rec;                // records written in this phase

    // one input byte; positions only ever move forward, so a chunk is loaded at most once
    __device__ __forceinline__ uint32_t byte(int32_t p)
    {
        const uint32_t q = (uint32_t) p + head;
        if ((q >> 5) != win_tag) {
            win_tag = q >> 5;
            const uint4 *src = reinterpret_cast<const uint4 *>(in - head + ((size_t) win_tag << 5));
            const uint4 a = __ldg(src), b = __ldg(src + 1);
            *reinterpret_cast<uint4 *>(win) = a;
            *reinterpret_cast<uint4 *>(win + 16) = b;
        }
        return win[q & 31];
    }
    // ll literals at input position lit_pos, then ml bytes copied from `off` back.  false: does not fit a record
    __device__ __forceinline__ bool emit(int32_t lit_pos, uint32_t ll, uint32_t ml, uint32_t off)
    {
        const int32_t skip = lit_pos - prev_lit_end;
        if (skip > kMaxSkip || ml > kMaxMatch || (off > kMaxOffset && off != kNoOffset)) return false;
        rec[n_rec++] = make_uint2(ll | (ml << 12), off | ((uint32_t) skip << 24));
        prev_lit_end = lit_pos + (int32_t) ll;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// execute: the records of one block (all lanes hold the same arguments)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void execute_records(const uint2 *rec, const int cnt, const uint8_t *in, uint8_t *out, uint32_t lit, uint32_t op, const int lane)
{
    for (int k = 0; k < cnt; k++) {
        const uint2 r = rec[k];
        const uint32_t ll = r.x & 0xfffu, ml = r.x >> 12, off = r.y & 0xffffffu;
        lit += r.y >> 24;
        const uint32_t total = ll + ml;
        if (total <= 32 && off >= total) {
            // the whole sequence in one step: every lane owns one output byte, a literal of the input or a match byte that
            // lies completely in front of this sequence (offset >= total)
            const uint32_t t = (uint32_t) lane;
            if (t < total) {
                const uint8_t *p = t < ll ? in + (lit + t) : out + (op + t - off);
                out[op + t] = *p;
            }
        }
        else {
            warp_copy(out + op, in + lit, ll, lane);
            if (ml) {
                __syncwarp();
                warp_match_copy(out + (op + ll), off, ml, lane);
            }
        }
        __syncwarp();
        lit += ll;
        op += total;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// one warp: claims blocks for its lanes, alternates parse and execute phases until the batch is exhausted
// ------------------------------------------------------------------------------------------------------------------
template <class Codec>
__device__ void run_warp(const AccBatch &b, WarpSmem &sm, const int lane, const int lanes_in_use)
{
    ParseCtx C;
    C.win = sm.win + lane * kWinStride;
    C.rec = sm.rec + lane * kRecStride;
    C.in = nullptr; C.in_len = 0; C.out_cap = 0; C.win_tag = ~0u; C.head = 0; C.prev_lit_end = 0; C.n_rec = 0;
    typename Codec::Parse P;
    Codec::begin(P);
    uint8_t *out = nullptr;
    uint32_t blk = 0;
    bool active = false, exhausted = lane >= lanes_in_use;
    for (;;) {
        // ---- claim ----
        if (!active && !exhausted) {
            const uint32_t idx = claim_block(b.work_counter);
            if ((int64_t) idx >= b.n) exhausted = true;
            else {
                blk = idx;
                const uint8_t *in = b.src + b.src_off[idx];
                const int64_t in_len = b.src_len[idx], out_cap = b.dst_cap[idx];
                out = b.dst + b.dst_off[idx];
                active = true;
                C.in = in;
                C.head = (uint32_t) ((uintptr_t) in & 31);
                C.win_tag = ~0u;
                C.prev_lit_end = 0;
                Codec::begin(P);
                if (in_len >= 0x7fffff00LL || out_cap >= 0x7fffff00LL || in_len < 32) { C.in_len = 0; C.out_cap = 0; Codec::whole(P); }
                else { C.in_len = (int32_t) in_len; C.out_cap = (int32_t) out_cap; }
            }
        }
        if (!__any_sync(kFull, active)) break;
        // ---- parse phase ----
        const uint32_t x_lit = (uint32_t) C.prev_lit_end, x_op = Codec::out_pos(P);
        C.n_rec = 0;
        bool fb = false;
        if (active) fb = Codec::is_whole(P) || Codec::parse_run(P, C, kRecPerLane) == kFallback;
        __syncwarp();
        // ---- execute phase ----
        unsigned todo = __ballot_sync(kFull, active && (C.n_rec > 0 || fb));
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int cnt = __shfl_sync(kFull, C.n_rec, src);
            const uint8_t *in_b = (const uint8_t *) __shfl_sync(kFull, (unsigned long long) C.in, src);
            uint8_t *out_b = (uint8_t *) __shfl_sync(kFull, (unsigned long long) out, src);
            const uint32_t lit_b = __shfl_sync(kFull, x_lit, src), op_b = __shfl_sync(kFull, x_op, src);
            if (cnt) execute_records(sm.rec + src * kRecStride, cnt, in_b, out_b, lit_b, op_b, lane);
            if (__shfl_sync(kFull, (int) fb, src)) {
                const uint32_t blk_b = __shfl_sync(kFull, blk, src);
                const uint32_t fip = __shfl_sync(kFull, Codec::fallback_ip(P), src), fop = __shfl_sync(kFull, Codec::fallback_op(P), src);
                Codec::finish(b, blk_b, fip, fop, lane);
                __syncwarp();
            }
        }
        if (fb) active = false;
    }
}

}  // namespace lzs
