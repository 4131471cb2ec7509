// lz_records.cuh -- the record path of the LZ4 and Snappy block decoders (sm_100a): the serial half and the wide half of
// an LZ decoder as two kernels with an 8-byte record per sequence in between.
//
//   PARSE KERNEL    one LANE per block: every lane walks the token grammar of its own block from the first token to the
//                   point where the reference decoder leaves its normal path (the end-of-block rules, malformed input,
//                   very long lengths), so a warp instruction advances 32 independent token chains and all blocks of
//                   the batch are parsed at the same time.  A lane reads its compressed bytes through a private ring of
//                   32-byte chunks in shared memory whose next chunk is always requested ahead of time (cp.async: no
//                   registers, no waiting; one global wavefront per ~4 sequences instead of one per byte) and writes RECORDS {literal length, match length, offset, header bytes
//                   skipped} to its block's row of a record table in global memory, then a header {record count, resume
//                   position}.
//   EXECUTE KERNEL  one WARP per block, from the first record to the last: 32 records per coalesced load, two warp
//                   prefix sums give every record's literal and output position (the lines are prefetched), then STEPS
//                   of up to 64 output bytes: as many consecutive records as fit and read their match bytes from in
//                   front of the step -- every lane resolves two bytes (which record: a population count over the
//                   record starts; then a literal of the input or an older output byte).  Output is built in a 4 KiB
//                   shared-memory ring per warp (matches nearer than that never touch L2; what a warp has just written is
//                   not in L1, stores do not allocate there) and leaves it in 16-byte stores.  Behind the last record the
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

constexpr int kWinChunks = 4;                   // per-lane input window of the parse kernel: a ring of four 32-byte chunks
constexpr int kWinBytes = 32 * kWinChunks;
constexpr int kWinStride = 144;                 // window pitch (16-byte aligned, spreads the lanes over the banks)
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
// 32 bytes global -> shared without passing through registers (LDGSTS); the lane goes on and waits only when it needs them
__device__ __forceinline__ void async_chunk(uint8_t *dst, const uint8_t *src)
{
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n\tcp.async.ca.shared.global [%2], [%3], 16;\n\tcp.async.commit_group;"
                 :: "r"(d), "l"(src), "r"(d + 16), "l"(src + 16) : "memory");
}
__device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
inline uint32_t claim_block(unsigned int *counter) { return __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED); }
inline void prefetch_l1(const void *) {}
inline void async_chunk(uint8_t *dst, const uint8_t *src) { memcpy(dst, src, 32); }
inline void async_wait() {}
#endif

// result of Codec::parse_one: one more sequence / element was recorded; the block leaves the fast path at the codec's resume
// point (kFallback), or its record row is full (kRowFull: same hand-over, the step decoder continues from the resume point)
enum ParseResult { kMore = 0, kRowFull = 1, kFallback = 2 };

// What a codec's parse loop sees of its lane's block.  All positions are relative to `in`.
struct ParseCtx {
    const uint8_t *in;        // first byte the grammar walk reads (behind the codec's preamble, if it has one)
    int32_t in_len, out_cap;
    uint8_t *win;             // this lane's window (shared memory): chunk c of the input lives at win + (c % 4) * 32
    uint2 *rec;               // this block's record row (global memory)
    uint32_t win_tag;         // chunks win_tag and win_tag + 1 are resident, win_tag + 2 is in flight or has landed; ~0: none
    uint32_t win_chunks;      // chunks that hold input at all (nothing is requested behind the block)
    uint32_t head;            // (address of in) & 31
    int32_t prev_lit_end;     // input position behind the literals of the previous record
    int n_rec;                // records written so far

    // Makes the 32 bytes from position p on readable without a check: the chunk of p and the one behind it are resident, the
    // third is requested (cp.async: no registers, no waiting).  Called once per sequence; in the common case (same chunk)
    // it is two instructions, crossing into the next chunk costs a wait that has usually been over for thousands of cycles --
    // the token chain of a lane is not stretched by memory latency, and neither are the 31 other chains of its warp.
    __device__ __forceinline__ void ensure(int32_t p)
    {
        const uint32_t c = ((uint32_t) p + head) >> 5;
        if (c == win_tag) return;
        const uint8_t *base = in - head;
        if (c == win_tag + 1 && win_tag != ~0u) {
            async_wait();                                 // chunk c + 1 (requested when c - 1 became current) has landed
        }
        else {                                            // a jump (long literal run) or the first chunk
            async_wait();                                 // nothing may still be in flight towards the slots that are reused
            async_chunk(win + (c % kWinChunks) * 32, base + ((size_t) c << 5));
            if (c + 1 < win_chunks) async_chunk(win + ((c + 1) % kWinChunks) * 32, base + ((size_t) (c + 1) << 5));
            async_wait();
        }
        win_tag = c;
        if (c + 2 < win_chunks) async_chunk(win + ((c + 2) % kWinChunks) * 32, base + ((size_t) (c + 2) << 5));
    }
    // one input byte at a position within 32 bytes behind the last ensure()
    __device__ __forceinline__ uint32_t byte(int32_t p) const { return win[((uint32_t) p + head) & (kWinBytes - 1)]; }
    // ll literals at input position lit_pos, then ml bytes copied from `off` back.  false: does not fit a record
    __device__ __forceinline__ bool emit(int32_t lit_pos, uint32_t ll, uint32_t ml, uint32_t off)
    {
        const int32_t skip = lit_pos - prev_lit_end;
        if (skip > kMaxSkip || ml > kMaxMatch || (off > kMaxOffset && off != kNoOffset)) return false;
        rec[n_rec++] = make_uint2(ll | (ml << 12), off | ((uint32_t) skip << 24));
        prev_lit_end = lit_pos + (int32_t) ll;
        return true;
    }
    // A sequence as one record {ll, ml, off}, or -- split -- as a literal-only record {ll} and a match record {ml, off} whose
    // (empty) literals sit at match_pos.  Written without a branch on `split`: the second slot is simply not counted when
    // it is not needed (the row has room for two: the callers check n_rec + 2 <= row).
    __device__ __forceinline__ bool emit2(int32_t lit_pos, uint32_t ll, uint32_t ml, uint32_t off, bool split, int32_t match_pos)
    {
        const int32_t skip = lit_pos - prev_lit_end;
        if (skip > kMaxSkip || ml > kMaxMatch || off > kMaxOffset || ll > (uint32_t) kMaxLitPiece || match_pos - lit_pos - (int32_t) ll > kMaxSkip) return false;
        rec[n_rec] = make_uint2(ll | (split ? 0u : ml << 12), (split ? kNoOffset : off) | ((uint32_t) skip << 24));
        rec[n_rec + 1] = make_uint2(ml << 12, off | ((uint32_t) (match_pos - lit_pos - (int32_t) ll) << 24));
        n_rec += split ? 2 : 1;
        prev_lit_end = split ? match_pos : lit_pos + (int32_t) ll;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// parse: one lane, block after block until the batch is exhausted.  `win` = this lane's window in shared memory.
// The 32 lanes of a warp run their rounds in lockstep (__syncwarp at the top of every round): a round then costs the
// instructions of the distinct paths its lanes take ONCE -- the common path is straight-line code for a whole sequence with
// up to one length-extension byte per length -- instead of letting 32 independent loops drift apart (measured: 380
// instructions per round with free-running lanes at 10.6 active lanes per instruction).
// ------------------------------------------------------------------------------------------------------------------
template <class Codec>
__device__ void parse_lane(const AccBatch &b, uint8_t *win, uint2 *recs, RecHeader *hdrs, const int row)
{
    bool have = false, done = false;
    uint32_t idx = 0;
    const uint8_t *in0 = nullptr;
    ParseCtx C;
    C.win = win; C.rec = nullptr; C.in = nullptr; C.in_len = 0; C.out_cap = 0; C.head = 0; C.win_chunks = 0; C.win_tag = ~0u; C.prev_lit_end = 0; C.n_rec = 0;
    typename Codec::Parse P;
    Codec::begin(P);
    for (;;) {
        __syncwarp();
        if (!have && !done) {
            idx = claim_block(b.work_counter);
            if ((int64_t) idx >= b.n) done = true;
            else {
                in0 = b.src + b.src_off[idx];
                const int64_t in_len = b.src_len[idx], out_cap = b.dst_cap[idx];
                if (in_len < 0x7fffff00LL && out_cap < 0x7fffff00LL && in_len >= 32) {
                    C.rec = recs + (size_t) idx * row;
                    C.in = in0; C.in_len = (int32_t) in_len; C.out_cap = (int32_t) out_cap;
                    C.head = (uint32_t) ((uintptr_t) in0 & 31);
                    C.win_chunks = (C.head + (uint32_t) in_len + 31) >> 5;
                    C.win_tag = ~0u; C.prev_lit_end = 0; C.n_rec = 0;
                    Codec::begin(P);
                    have = true;
                }
                else {                                    // tiny or huge block: the step decoder takes all of it
                    RecHeader h;
                    h.n_rec = 0; h.resume_ip = kWholeBlock; h.resume_op = 0; h.preamble = 0;
                    hdrs[idx] = h;
                }
            }
        }
        if (__all_sync(kFull, done)) return;
        if (have && Codec::parse_one(P, C, row) != kMore) {
            RecHeader h;
            h.n_rec = (uint32_t) C.n_rec;
            h.resume_ip = Codec::resume_ip(P);
            h.resume_op = Codec::resume_op(P);
            h.preamble = (uint32_t) (C.in - in0);
            hdrs[idx] = h;
            have = false;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// execute: one warp, block after block
// ------------------------------------------------------------------------------------------------------------------
constexpr int kOutRing = 4096;                  // decoded bytes a warp keeps in shared memory (matches nearer than this never touch L2)
constexpr int kFlushBytes = 512;                // output leaves the ring in pieces of this size (32 lanes x 16 bytes)
constexpr uint32_t kRingMask = kOutRing - 1;

// Moves [flushed, e) from the ring to global memory: 16-byte stores for whole aligned units, bytes for the partial unit at
// the start of a block and (final flush) at its end.  Positions are output position + (output address & 15), so ring
// units and global units are aligned alike.
__device__ __forceinline__ void flush_to(const uint8_t *ring, uint8_t *out_al, uint32_t &flushed, const uint32_t e, const int lane)
{
    uint32_t a = flushed;
    if (a & 15u) {
        uint32_t a1 = (a + 15u) & ~15u;
        if (a1 > e) a1 = e;
        if (a + (uint32_t) lane < a1) out_al[a + lane] = ring[(a + lane) & kRingMask];
        a = a1;
    }
    const uint32_t units = (e - a) >> 4;
    for (uint32_t u = (uint32_t) lane; u < units; u += 32) {
        const uint32_t w = a + (u << 4);
        *reinterpret_cast<uint4 *>(out_al + w) = *reinterpret_cast<const uint4 *>(ring + (w & kRingMask));
    }
    a += units << 4;
    if (a + (uint32_t) lane < e) out_al[a + lane] = ring[(a + lane) & kRingMask];
    flushed = e;
    __syncwarp();
}

// one record the multi-record step cannot take (more than 64 bytes, or a match that overlaps its own output): literals, then
// the match in pieces of at most kFlushBytes so that the ring can drain in between
__device__ __forceinline__ void execute_long_record(uint8_t *ring, const uint8_t *in, uint8_t *out_al, uint32_t &flushed, const uint32_t litp,
                                                    const uint32_t opw, const uint32_t ll, const uint32_t ml, const uint32_t off, const int lane)
{
    for (uint32_t cb = 0; cb < ll; cb += kFlushBytes) {
        const uint32_t pn = ll - cb < (uint32_t) kFlushBytes ? ll - cb : (uint32_t) kFlushBytes;
        for (uint32_t i = (uint32_t) lane; i < pn; i += 32) ring[(opw + cb + i) & kRingMask] = __ldg(in + (litp + cb + i));
        __syncwarp();
        if (opw + cb + pn - flushed >= (uint32_t) kFlushBytes) flush_to(ring, out_al, flushed, (opw + cb + pn) & ~15u, lane);
    }
    const uint32_t mopw = opw + ll;
    for (uint32_t cb = 0; cb < ml; cb += kFlushBytes) {
        const uint32_t pw = mopw + cb;                              // first byte of this piece
        const uint32_t pn = ml - cb < (uint32_t) kFlushBytes ? ml - cb : (uint32_t) kFlushBytes;
        if (off >= 32) {
            for (uint32_t base = 0; base < pn; base += 32) {        // a round only reads bytes written at least 32 positions earlier
                const uint32_t i = base + (uint32_t) lane;
                if (i < pn) {
                    const uint32_t pos = pw + i, src = pos - off;
                    uint32_t v;
                    if ((int32_t) (src - (pw + base + 32 - kOutRing)) >= 0) v = ring[src & kRingMask];
                    else v = out_al[src];                           // older than the ring: flushed long ago
                    ring[pos & kRingMask] = (uint8_t) v;
                }
                __syncwarp();
            }
        }
        else {
            // periodic pattern: every byte of the piece repeats one of the `off` bytes in front of it
            uint32_t m = (uint32_t) lane % off;
            const uint32_t step = 32u % off;
            for (uint32_t i = (uint32_t) lane; i < pn; i += 32) {
                ring[(pw + i) & kRingMask] = ring[(pw - off + m) & kRingMask];
                m += step;
                if (m >= off) m -= off;
            }
            __syncwarp();
        }
        if (pw + pn - flushed >= (uint32_t) kFlushBytes) flush_to(ring, out_al, flushed, (pw + pn) & ~15u, lane);
    }
}

template <class Codec>
__device__ void execute_warp(const AccBatch &b, const uint2 *recs, const RecHeader *hdrs, const int row, uint8_t *ring, const int lane)
{
    const uint32_t le_mask = 0xffffffffu >> (31 - lane);             // lanes <= mine
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = claim_block(b.work_counter);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) return;
        const RecHeader h = hdrs[idx];
        if (h.n_rec) {
            const uint8_t *in = b.src + b.src_off[idx] + h.preamble;
            uint8_t *out = b.dst + b.dst_off[idx];
            const uint32_t oh = (uint32_t) ((uintptr_t) out & 15);
            uint8_t *out_al = out - oh;                               // positions below are output position + oh
            const uint2 *rec = recs + (size_t) idx * row;
            uint32_t flushed = oh;
            uint32_t lit = 0, opw = oh;                               // running positions at the start of the batch (uniform)
            for (uint32_t base = 0; base < h.n_rec; base += 32) {
                const uint32_t cnt = h.n_rec - base < 32u ? h.n_rec - base : 32u;
                uint2 mine = make_uint2(0, 0);
                if ((uint32_t) lane < cnt) mine = rec[base + lane];
                const uint32_t my_ll = mine.x & 0xfffu, my_ml = mine.x >> 12, my_off = mine.y & 0xffffffu, my_total = my_ll + my_ml;
                // where the literals and the output of MY record lie: prefix sums over the batch
                uint32_t a = (mine.y >> 24) + my_ll, t = my_total;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t ua = __shfl_up_sync(kFull, a, o), ut = __shfl_up_sync(kFull, t, o);
                    if (lane >= o) { a += ua; t += ut; }
                }
                const uint32_t my_lit = lit + a - my_ll;              // input position of my literals
                const uint32_t my_opw = opw + t - my_total;           // where my output starts
                const uint32_t my_mop = my_opw + my_ll;               // ... and where my match starts
                const uint32_t my_dl = my_lit - my_opw;               // literal address = output position + my_dl
                if ((uint32_t) lane < cnt) {                          // ask for the lines now: the steps below then find them in L1
                    if (my_ll) prefetch_l1(in + my_lit);
                    if (my_ml && my_off <= my_mop && my_mop - my_off + kOutRing < my_mop + my_ml) prefetch_l1(out_al + (my_mop - my_off));
                }
                uint32_t f = 0;
                while (f < cnt) {
                    // ---- a step: as many consecutive records as end within 64 bytes of the first one's start S and take all
                    // their match bytes from in front of S (then no byte of the step depends on another byte of the step)
                    const uint32_t S = __shfl_sync(kFull, my_opw, (int) f);
                    const uint32_t endk = my_opw + my_total - S;
                    const bool fits = (uint32_t) lane >= f && (uint32_t) lane < cnt && endk <= 64 && (my_ml == 0 || my_off >= endk);
                    const uint32_t run = __ballot_sync(kFull, fits) >> f;
                    const uint32_t nfit = (uint32_t) __ffs((int) ~run) - 1;
                    if (nfit == 0) {
                        execute_long_record(ring, in, out_al, flushed, __shfl_sync(kFull, my_lit, (int) f), S, __shfl_sync(kFull, my_ll, (int) f),
                                            __shfl_sync(kFull, my_ml, (int) f), __shfl_sync(kFull, my_off, (int) f), lane);
                        f++;
                        continue;
                    }
                    const uint32_t E = __shfl_sync(kFull, endk, (int) (f + nfit - 1));   // bytes of this step
                    // which record produces byte j of the step: count the record starts at or below j
                    const bool inwin = (uint32_t) lane >= f && (uint32_t) lane < f + nfit;
                    const uint32_t st = my_opw - S;
                    const uint32_t lo = __reduce_or_sync(kFull, (inwin && st < 32) ? 1u << st : 0u);
                    {
                        const uint32_t r = f - 1 + (uint32_t) __popc(lo & le_mask);
                        const uint32_t mop = __shfl_sync(kFull, my_mop, (int) r), dl = __shfl_sync(kFull, my_dl, (int) r), off = __shfl_sync(kFull, my_off, (int) r);
                        const uint32_t pos = S + (uint32_t) lane;
                        if ((uint32_t) lane < E) {
                            uint32_t v;
                            if (pos < mop) v = __ldg(in + (pos + dl));
                            else {
                                const uint32_t src = pos - off;
                                v = (int32_t) (src - (S + E - kOutRing)) >= 0 ? ring[src & kRingMask] : out_al[src];
                            }
                            ring[pos & kRingMask] = (uint8_t) v;
                        }
                    }
                    if (E > 32) {
                        const uint32_t hi = __reduce_or_sync(kFull, (inwin && st >= 32) ? 1u << (st - 32) : 0u);
                        const uint32_t r = f - 1 + (uint32_t) __popc(lo) + (uint32_t) __popc(hi & le_mask);
                        const uint32_t mop = __shfl_sync(kFull, my_mop, (int) r), dl = __shfl_sync(kFull, my_dl, (int) r), off = __shfl_sync(kFull, my_off, (int) r);
                        const uint32_t pos = S + 32 + (uint32_t) lane;
                        if ((uint32_t) lane + 32 < E) {
                            uint32_t v;
                            if (pos < mop) v = __ldg(in + (pos + dl));
                            else {
                                const uint32_t src = pos - off;
                                v = (int32_t) (src - (S + E - kOutRing)) >= 0 ? ring[src & kRingMask] : out_al[src];
                            }
                            ring[pos & kRingMask] = (uint8_t) v;
                        }
                    }
                    __syncwarp();
                    f += nfit;
                    if (S + E - flushed >= (uint32_t) kFlushBytes) flush_to(ring, out_al, flushed, (S + E) & ~15u, lane);
                }
                lit += __shfl_sync(kFull, a, 31);
                opw += __shfl_sync(kFull, t, 31);
            }
            if (opw != flushed) flush_to(ring, out_al, flushed, opw, lane);   // everything in front of the resume point is in global memory now
        }
        // the step decoder finishes the block (its tail at least) from the recorded position; writes out_len / status
        Codec::finish(b, idx, h.resume_ip, h.resume_op, lane);
        __syncwarp();
    }
}

}  // namespace lzs
