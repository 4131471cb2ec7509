// lz_records.cuh -- the record path of the LZ4 and Snappy block decoders (sm_100a): the serial half and the wide half of
// an LZ decoder as two kernels with an 8-byte record per sequence in between.
//
//   PARSE KERNEL    one LANE per block: every lane walks the token grammar of its own block from the first token to the
//                   point where the reference decoder leaves its normal path (the end-of-block rules, malformed input,
//                   very long lengths), so a warp instruction advances 32 independent token chains and all blocks of
//                   the batch are parsed at the same time.  A lane reads its compressed bytes through a private 32-byte
//                   shared-memory window (two aligned 16-byte loads per refill: one global wavefront per ~4 sequences
//                   instead of one per byte) and writes RECORDS {literal length, match length, offset, header bytes
//                   skipped} to its block's row of a record table in global memory, then a header {record count, resume
//                   position}.
//   EXECUTE KERNEL  one WARP per block, from the first record to the last: 32 records per coalesced load, two warp
//                   prefix sums give every record's literal and output position (the lines are prefetched: older output
//                   is not in L1, stores do not allocate there), then every short sequence is one load (a literal of
//                   the input or an older output byte, selected per lane) and one store per lane; long literal runs and
//                   long / overlapping matches use the 16-byte warp copies of acc_device.cuh.  Behind the last record the
//                   warp resumes the step decoder (lz4_decode_v1.cuh / snappy_decode.cuh) at the recorded position: that
//                   is the exact restatement of the Java loop, so accept/reject decisions, error offsets and the bytes
//                   of the block tail are those of the reference by construction.  A block with more sequences than its
//                   row holds simply resumes there earlier.
//
// The parse lanes only accept what they can prove the reference decoder takes on its normal path
// (Lz4RawDecompressor.java:59-195, SnappyRawDecompressor.java:70-220).
//
// History: two fused versions of this split lost to the step decoder -- a parse warp feeding execute warps
// through shared-memory queues (one warp cannot issue fast enough to parse for an SM), and warps alternating between
// parsing 32 blocks and executing them (all output windows of the batch live at once: every match source came from
// DRAM).  Two kernels keep what was right in both: parsing costs one warp instruction per 32 sequences, and only one
// block per warp is being written at any time.
//
// The same source compiles for the host with LZS_EMU defined (tests/host/lzs_emu.cpp: OS threads as lanes), which is how
// the parse logic and the record hand-over are checked on the CPU (tests/test_record_engine_emu.py).
#pragma once
#include "acc_device.cuh"

namespace lzs {

constexpr int kWinBytes = 32;                   // per-lane input window of the parse kernel
constexpr int kWinStride = 48;                  // window pitch (16-byte aligned, spreads the lanes over the banks)
constexpr uint32_t kNoOffset = 0xffffffu;       // offset field of a literal-only record
constexpr int kMaxLitPiece = 4095;              // ll field: 12 bits (longer runs are cut into pieces)
constexpr uint32_t kMaxMatch = (1u << 20) - 1;  // ml field: 20 bits
constexpr uint32_t kMaxOffset = (1u << 24) - 2; // off field: 24 bits
constexpr int kMaxSkip = 255;                   // skip field: 8 bits
constexpr uint32_t kWholeBlock = 0xffffffffu;   // resume position meaning "the step decoder decodes the whole block"

struct __align__(16) RecHeader {
    uint32_t n_rec;       // records of this block
    uint32_t resume_ip;   // where the step decoder takes over (input position behind the codec's preamble), or kWholeBlock
    uint32_t resume_op;   // ... and the output position there
    uint32_t preamble;    // bytes of the codec's preamble in front of position 0 (Snappy's length varint)
};

#ifndef LZS_EMU
__device__ __forceinline__ uint32_t claim_block(unsigned int *counter) { return atomicAdd(counter, 1u); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#else
inline uint32_t claim_block(unsigned int *counter) { return __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED); }
inline void prefetch_l1(const void *) {}
#endif

enum ParseResult { kRowFull = 0, kFallback = 2 };

// What a codec's parse loop sees of its lane's block.  All positions are relative to `in`.
struct ParseCtx {
    const uint8_t *in;        // first byte the grammar walk reads (behind the codec's preamble, if it has one)
    int32_t in_len, out_cap;
    uint8_t *win;             // this lane's 32-byte window (shared memory)
    uint2 *rec;               // this block's record row (global memory)
    uint32_t win_tag;         // the aligned 32-byte chunk the window holds; ~0: none
    uint32_t head;            // (address of in) & 31
    int32_t prev_lit_end;     // input position behind the literals of the previous record
    int n_rec;                // records written so far

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
// parse: one lane, block after block until the batch is exhausted.  `win` = this lane's window in shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <class Codec>
__device__ void parse_lane(const AccBatch &b, uint8_t *win, uint2 *recs, RecHeader *hdrs, const int row)
{
    for (;;) {
        const uint32_t idx = claim_block(b.work_counter);
        if ((int64_t) idx >= b.n) return;
        const uint8_t *in = b.src + b.src_off[idx];
        const int64_t in_len = b.src_len[idx], out_cap = b.dst_cap[idx];
        RecHeader h;
        h.n_rec = 0; h.resume_ip = kWholeBlock; h.resume_op = 0; h.preamble = 0;
        if (in_len < 0x7fffff00LL && out_cap < 0x7fffff00LL && in_len >= 32) {
            ParseCtx C;
            C.win = win;
            C.rec = recs + (size_t) idx * row;
            C.in = in; C.in_len = (int32_t) in_len; C.out_cap = (int32_t) out_cap;
            C.head = (uint32_t) ((uintptr_t) in & 31);
            C.win_tag = ~0u; C.prev_lit_end = 0; C.n_rec = 0;
            typename Codec::Parse P;
            Codec::begin(P);
            Codec::parse_run(P, C, row);                  // until the block leaves the fast path or the row is full
            h.n_rec = (uint32_t) C.n_rec;
            h.resume_ip = Codec::resume_ip(P);
            h.resume_op = Codec::resume_op(P);
            h.preamble = (uint32_t) (C.in - in);
        }
        hdrs[idx] = h;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// execute: one warp, block after block
// ------------------------------------------------------------------------------------------------------------------
template <class Codec>
__device__ void execute_warp(const AccBatch &b, const uint2 *recs, const RecHeader *hdrs, const int row, const int lane)
{
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = claim_block(b.work_counter);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) return;
        const RecHeader h = hdrs[idx];
        if (h.n_rec) {
            const uint8_t *in = b.src + b.src_off[idx] + h.preamble;
            uint8_t *out = b.dst + b.dst_off[idx];
            const uint2 *rec = recs + (size_t) idx * row;
            uint32_t lit = 0, op = 0;                                 // running positions (all lanes hold the same values)
            for (uint32_t base = 0; base < h.n_rec; base += 32) {
                const uint32_t cnt = h.n_rec - base < 32u ? h.n_rec - base : 32u;
                uint2 mine = make_uint2(0, 0);
                if ((uint32_t) lane < cnt) mine = rec[base + lane];
                // where the literals and the match source of MY record lie: prefix sums over the batch; ask for the lines now
                {
                    const uint32_t my_ll = mine.x & 0xfffu, my_ml = mine.x >> 12, my_off = mine.y & 0xffffffu;
                    uint32_t a = (mine.y >> 24) + my_ll, t = my_ll + my_ml;
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t ua = __shfl_up_sync(kFull, a, o), ut = __shfl_up_sync(kFull, t, o);
                        if (lane >= o) { a += ua; t += ut; }
                    }
                    if ((uint32_t) lane < cnt) {
                        if (my_ll) prefetch_l1(in + (lit + a - my_ll));
                        const uint32_t mop = op + t - my_ml;                  // output position of my match
                        if (my_ml && my_off <= mop && mop - my_off < op) prefetch_l1(out + (mop - my_off));   // only what is already written
                    }
                }
                for (uint32_t k = 0; k < cnt; k++) {
                    const uint32_t w0 = __shfl_sync(kFull, mine.x, (int) k), w1 = __shfl_sync(kFull, mine.y, (int) k);
                    const uint32_t ll = w0 & 0xfffu, ml = w0 >> 12, off = w1 & 0xffffffu;
                    lit += w1 >> 24;
                    const uint32_t total = ll + ml;
                    if (total <= 32 && off >= total) {
                        // the whole sequence in one step: every lane owns one output byte, a literal of the input or a match
                        // byte that lies completely in front of this sequence (offset >= total)
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
        }
        // the step decoder finishes the block (its tail at least) from the recorded position; writes out_len / status
        Codec::finish(b, idx, h.resume_ip, h.resume_op, lane);
        __syncwarp();
    }
}

}  // namespace lzs
