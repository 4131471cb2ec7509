// lz_stream.cuh -- the streaming LZ77 decode engine shared by the LZ4 and Snappy block decoders (sm_100a).
//
// A byte-oriented LZ decoder has two very different halves: walking the token grammar is a serial chain of tiny
// dependent steps (one token decides where the next one starts), while producing the bytes is wide copy work.  This
// engine gives each half the execution shape it wants:
//
//   * one PARSE warp per CTA walks the grammars of up to kSlots blocks at once, ONE BLOCK PER LANE, so a warp
//     instruction advances up to 32 independent token chains.  A lane reads its block's compressed bytes from a
//     2 KiB shared-memory ring that it keeps filled with 512-byte cp.async.bulk copies (mbarrier complete_tx), and
//     turns sequences / elements into 16-byte RECORDS {literal position, output position, literal length, match
//     length, offset} in a per-block shared-memory queue;
//   * one EXECUTE warp per block drains that queue: literals come from the input ring, match bytes from a 4 KiB
//     shared-memory output ring (only matches farther back than the ring go to L2), a whole short sequence is one
//     shared-memory load + store per lane, and finished output leaves the ring in 16-byte stores (512 B per warp
//     instruction pair).
//
// The parse lanes only accept what they can prove the reference decoder would take on its normal path
// (Lz4RawDecompressor.java:59-195, SnappyRawDecompressor.java:70-220); everything else -- the end-of-block rules,
// malformed input, very long lengths -- ends the fast path with a FALLBACK record at a token boundary, and the execute
// warp finishes the block from that (input, output) position with the exact restatement of the Java loop
// (lz4_decode_v1.cuh general path / snappy_decode_from).  Accept/reject decisions, error offsets and output bytes are
// therefore those of the general path by construction.
//
// The same source compiles for the host with LZS_EMU defined (tests/host/lzs_emu.cpp: OS threads as lanes), which is
// how the queue / ring protocol is checked on the CPU before it ever runs on a GPU.
#pragma once
#include "acc_device.cuh"

namespace lzs {

constexpr int kInRing = 2048;             // compressed bytes staged per block slot
constexpr int kChunk = 512;               // one bulk copy
constexpr int kNChunk = kInRing / kChunk;
constexpr int kOutRing = 4096;            // decoded bytes kept in shared memory per block slot
constexpr int kNRec = 64;                 // record queue entries per block slot
constexpr int kBatch = 16;                // records an execute warp takes before it publishes its progress
constexpr int kLitPiece = 1024;           // longest literal-only record (long runs are cut into pieces)
constexpr int kFlushBytes = 512;          // output leaves the ring in pieces of this size (32 lanes x 16 bytes)
constexpr uint32_t kNoOffset = 0x7fffffffu;
constexpr uint32_t kSpinLimit = 1u << 24;  // polls without progress before a wait is declared dead (seconds; a launch takes milliseconds)

// control records have z == 0; w says which
constexpr uint32_t kRecBegin = 1;         // x = block index
constexpr uint32_t kRecFallback = 2;      // x = input position, y = output position (block space): finish with the general path
constexpr uint32_t kRecExit = 3;
constexpr uint32_t kFallbackWhole = 0xffffffffu;   // x of a FALLBACK record: decode the whole block with the general path (preamble included)

struct __align__(16) Slot {
    uint8_t in_ring[kInRing];
    uint8_t out_ring[kOutRing];
    uint4 rec[kNRec];
    unsigned long long mbar[kNChunk];
    uint32_t prod;        // records produced (parse lane)
    uint32_t cons;        // records consumed (execute warp)
    uint32_t cons_q;      // running input position below which the execute warp needs nothing any more
    uint32_t abort;       // watchdog: set when a wait took implausibly long; everybody leaves
};
static_assert(sizeof(Slot) % 16 == 0, "slot alignment");

// ------------------------------------------------------------------------------------------------------------------
// platform layer: PTX on the device, plain atomics in the host emulation
// ------------------------------------------------------------------------------------------------------------------
#ifndef LZS_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p) { uint32_t v; asm volatile("ld.relaxed.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long *b) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one bulk copy global -> shared, completion (byte count) signalled on `bar`; src/dst 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t claim_block(unsigned int *counter) { return atomicAdd(counter, 1u); }
__device__ __forceinline__ void backoff(unsigned ns) { __nanosleep(ns); }
__device__ __forceinline__ bool any_lane(bool v) { return __any_sync(__activemask(), v); }   // a scheduling hint only
__device__ __forceinline__ uint8_t ld_far(const uint8_t *p) { return __ldcg(p); }   // older output of this block: L2 (bypasses L1, always coherent)
// a shared-memory word read by all lanes of a warp in one instruction: every lane sees the same value
__device__ __forceinline__ uint32_t warp_ld_acquire(const uint32_t *p, int) { return ld_acquire(p); }
__device__ __forceinline__ uint32_t warp_ld_relaxed(const uint32_t *p, int) { return ld_relaxed(p); }
#else
// host emulation (tests/host/lzs_emu.cpp): threads as lanes, a DMA thread lands the bulk copies late and out of order
inline void st_release(uint32_t *p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline uint32_t ld_acquire(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline uint32_t ld_relaxed(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void st_relaxed(uint32_t *p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline void mbar_init(unsigned long long *b) { *b = 0; }
inline void fence_proxy_async() {}
void emu_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar);
void emu_count_record(uint32_t z, uint32_t w, uint32_t x, uint32_t y);
inline void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) { emu_bulk_load(dst, src, bytes, bar); }
inline bool mbar_test(unsigned long long *b, uint32_t parity) { return ((uint32_t) __atomic_load_n(b, __ATOMIC_ACQUIRE) & 1u) != parity; }
inline uint32_t claim_block(unsigned int *counter) { return __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED); }
inline void backoff(unsigned) { sched_yield(); }
inline bool any_lane(bool v) { return v; }
inline uint8_t ld_far(const uint8_t *p) { return *(const volatile uint8_t *) p; }
inline uint32_t warp_bcast(uint32_t v, int lane)
{
    if (lane == 0) t_warp->bcast = v;
    __syncwarp();
    const uint32_t r = t_warp->bcast;
    __syncwarp();
    return r;
}
inline uint32_t warp_ld_acquire(const uint32_t *p, int lane) { return warp_bcast(lane == 0 ? ld_acquire(p) : 0, lane); }
inline uint32_t warp_ld_relaxed(const uint32_t *p, int lane) { return warp_bcast(lane == 0 ? ld_relaxed(p) : 0, lane); }
#endif

struct BlockDesc {
    const uint8_t *in;
    uint8_t *out;
    int64_t in_len, out_cap;
};
__device__ __forceinline__ BlockDesc load_desc(const AccBatch &b, uint32_t idx)
{
    BlockDesc d;
    d.in = b.src + b.src_off[idx];
    d.in_len = b.src_len[idx];
    d.out = b.dst + b.dst_off[idx];
    d.out_cap = b.dst_cap[idx];
    return d;
}

// ------------------------------------------------------------------------------------------------------------------
// parse side
// ------------------------------------------------------------------------------------------------------------------
enum ParseResult { kProgress = 0, kWait = 1, kFallback = 2 };

// what a codec's parse step sees of its lane's block
struct ParseCtx {
    Slot *S;
    uint32_t Qb;          // running ring position of block position 0
    int32_t in_len;       // block length (after the codec's preamble, if it has one)
    int32_t out_cap;
    int32_t avail;        // block positions [0, avail) have landed in the ring
    uint32_t oh;          // output misalignment: output position p lives at ring / flush position p + oh
    uint32_t prod;        // records written so far (published by the engine at the end of the round)

    __device__ __forceinline__ uint32_t byte(int32_t p) const { return S->in_ring[(Qb + (uint32_t) p) & (kInRing - 1)]; }
    __device__ __forceinline__ void emit(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
    {
        S->rec[prod & (kNRec - 1)] = make_uint4(x, y, z, w);
        prod++;
#ifdef LZS_EMU
        emu_count_record(z, w, x, y);
#endif
    }
    // a sequence: ll literals at block position lit_pos, then ml bytes copied from `off` back; output starts at op
    __device__ __forceinline__ void emit_seq(int32_t lit_pos, int32_t op, uint32_t ll, uint32_t ml, uint32_t off)
    {
        emit(Qb + (uint32_t) lit_pos, (uint32_t) op + oh, ll | (ml << 12), off);
    }
};

// One lane of the parse warp: claims blocks, keeps the input ring filled, runs the codec's parse step.
template <class Codec>
__device__ void parse_lane(const AccBatch &b, Slot &S)
{
    enum { kIdle = 0, kRun = 1 };
    int state = kIdle;
    uint32_t g_issued = 0, g_ready = 0, g_end = 0, g0 = 0;   // running chunk counters (never reset: they carry the mbarrier phases)
    const uint8_t *in_al = nullptr;                          // 16-byte aligned start of the first chunk
    uint32_t stream_q = 0;                                   // head + block length: bytes of aligned stream
    uint32_t published = 0;
    uint32_t spins = 0;
    ParseCtx C;
    C.S = &S; C.prod = 0; C.Qb = 0; C.in_len = 0; C.out_cap = 0; C.avail = 0; C.oh = 0;
    typename Codec::Parse P;
    for (;;) {
        const uint32_t cons = ld_acquire(&S.cons);
        bool progress = false;
        // ---- input ring: issue the next chunk when its ring slot is free, note chunks that have landed ----
        if (g_issued != g_end) {
            const uint32_t cq = ld_relaxed(&S.cons_q);
            if ((int32_t) (cq - (g_issued - (kNChunk - 1)) * kChunk) >= 0) {
                const uint32_t c = g_issued - g0;
                uint32_t bytes = stream_q - c * kChunk;
                bytes = bytes >= (uint32_t) kChunk ? (uint32_t) kChunk : ((bytes + 15u) & ~15u);
                fence_proxy_async();    // the ring slot was last read through the generic proxy
                bulk_load(S.in_ring + (g_issued & (kNChunk - 1)) * kChunk, in_al + (size_t) c * kChunk, bytes, &S.mbar[g_issued & (kNChunk - 1)]);
                g_issued++;
                progress = true;
            }
        }
        if (g_ready != g_issued && mbar_test(&S.mbar[g_ready & (kNChunk - 1)], (g_ready / kNChunk) & 1)) {
            g_ready++;
            progress = true;
        }
        if (state == kIdle) {
            // the previous block is finished when the execute warp has taken its FALLBACK record and every bulk copy
            // issued for it has landed
            if (cons == C.prod && g_ready == g_issued) {
                const uint32_t idx = claim_block(b.work_counter);
                if ((int64_t) idx >= b.n) {
                    C.emit(0, 0, 0, kRecExit);
                    st_release(&S.prod, C.prod);
                    return;
                }
                const BlockDesc d = load_desc(b, idx);
                C.emit(idx, 0, 0, kRecBegin);
                const bool big = d.in_len >= 0x7fffff00LL || d.out_cap >= 0x7fffff00LL;
                if (big || d.in_len < 32) {
                    C.emit(kFallbackWhole, 0, 0, kRecFallback);
                }
                else {
                    const uint32_t head = (uint32_t) ((uintptr_t) d.in & 15);
                    in_al = d.in - head;
                    stream_q = head + (uint32_t) d.in_len;
                    g0 = g_issued;
                    g_end = g0 + (stream_q + kChunk - 1) / kChunk;
                    C.Qb = g0 * kChunk + head;
                    C.in_len = (int32_t) d.in_len;
                    C.out_cap = (int32_t) d.out_cap;
                    C.oh = (uint32_t) ((uintptr_t) d.out & 15);
                    st_relaxed(&S.cons_q, g0 * kChunk);   // the execute warp is idle: nobody else writes this now
                    Codec::begin(P);
                    state = kRun;
                }
                progress = true;
            }
        }
        else if (C.prod - cons < (uint32_t) kNRec) {
            const int32_t landed = (int32_t) ((g_ready - g0) * kChunk) - (int32_t) (C.Qb - g0 * kChunk);
            C.avail = landed < C.in_len ? landed : C.in_len;
            const int r = Codec::parse_step(P, C);
            if (r == kFallback) {
                C.emit(Codec::fallback_ip(P), Codec::fallback_op(P), 0, kRecFallback);
                g_end = g_issued;       // no further chunks of this block
                state = kIdle;
                progress = true;
            }
            else if (r == kProgress) progress = true;
        }
        if (C.prod != published) {
            st_release(&S.prod, C.prod);
            published = C.prod;
        }
        if (progress) spins = 0;
        else if (++spins > kSpinLimit || ld_relaxed(&S.abort)) { st_relaxed(&S.abort, 1); return; }
        if (!any_lane(progress)) backoff(64);   // nothing to do for any block of this warp: leave the issue slots to the execute warps
    }
}

// ------------------------------------------------------------------------------------------------------------------
// execute side
// ------------------------------------------------------------------------------------------------------------------
// Moves [flushed, e) from the output ring to global memory: 16-byte stores for whole aligned units, bytes for the partial
// unit at the start of a block and (final flush only) at its end.  Positions are block output position + oh.
__device__ __forceinline__ void flush_to(const uint8_t *ring, uint8_t *out_al, uint32_t &flushed, const uint32_t e, const int lane)
{
    uint32_t a = flushed;
    if (a & 15u) {
        uint32_t a1 = (a + 15u) & ~15u;
        if (a1 > e) a1 = e;
        if (a + (uint32_t) lane < a1) out_al[a + lane] = ring[(a + lane) & (kOutRing - 1)];
        a = a1;
    }
    const uint32_t units = (e - a) >> 4;
    for (uint32_t u = (uint32_t) lane; u < units; u += 32) {
        const uint32_t w = a + (u << 4);
        *reinterpret_cast<uint4 *>(out_al + w) = *reinterpret_cast<const uint4 *>(ring + (w & (kOutRing - 1)));
    }
    a += units << 4;
    if (a + (uint32_t) lane < e) out_al[a + lane] = ring[(a + lane) & (kOutRing - 1)];
    flushed = e;
    __syncwarp();
}

template <class Codec>
__device__ void exec_warp(const AccBatch &b, Slot &S, const int lane)
{
    uint8_t *const sb = S.in_ring;               // in_ring at [0, kInRing), out_ring behind it
    uint8_t *const ring = S.out_ring;
    uint32_t cons = 0;
    uint32_t blk = 0, flushed = 0;
    uint8_t *out_al = nullptr;
    uint32_t spins = 0;
    for (;;) {
        const uint32_t prod = warp_ld_acquire(&S.prod, lane);
        if (prod == cons) {
            backoff(64);
            if (++spins > kSpinLimit || warp_ld_relaxed(&S.abort, lane)) {
                if (lane == 0) { st_relaxed(&S.abort, 1); if (out_al) { b.out_len[blk] = 0; b.status[blk] = ACC_STATUS(ACC_E_CUDA, 0); } }
                return;
            }
            continue;
        }
        spins = 0;
        uint32_t n = prod - cons;
        if (n > (uint32_t) kBatch) n = kBatch;
        uint32_t last_q = 0;
        bool have_q = false;
        for (uint32_t k = 0; k < n; k++) {
            const uint4 r = S.rec[(cons + k) & (kNRec - 1)];
            if (r.z != 0) {
                const uint32_t ll = r.z & 0xfffu, ml = r.z >> 12, off = r.w, opw = r.y, lq = r.x;
                const uint32_t total = ll + ml;
                const uint32_t endw = opw + total;
                last_q = lq + ll;
                have_q = true;
                if (total <= 32 && off >= total) {
                    // the whole sequence in one step: every lane owns one output byte, a literal from the input ring or a
                    // match byte that lies completely in front of this sequence (offset >= total)
                    const uint32_t t = (uint32_t) lane;
                    if (t < total) {
                        const uint32_t srcw = opw + t - off;
                        const bool lit = t < ll;
                        uint32_t v;
                        if (lit || (int32_t) (srcw - (endw - kOutRing)) >= 0) v = sb[lit ? ((lq + t) & (kInRing - 1)) : (kInRing + (srcw & (kOutRing - 1)))];
                        else v = ld_far(out_al + srcw);
                        ring[(opw + t) & (kOutRing - 1)] = (uint8_t) v;
                    }
                    __syncwarp();
                }
                else {
                    for (uint32_t i = (uint32_t) lane; i < ll; i += 32) ring[(opw + i) & (kOutRing - 1)] = sb[(lq + i) & (kInRing - 1)];
                    __syncwarp();
                    const uint32_t mopw = opw + ll;
                    // matches run in pieces of at most kFlushBytes so that the ring can drain in between
                    for (uint32_t cb = 0; cb < ml; cb += kFlushBytes) {
                        const uint32_t pw = mopw + cb;                              // first byte of this piece
                        const uint32_t pn = ml - cb < (uint32_t) kFlushBytes ? ml - cb : (uint32_t) kFlushBytes;
                        if (off >= 32) {
                            for (uint32_t base = 0; base < pn; base += 32) {
                                const uint32_t i = base + (uint32_t) lane;
                                if (i < pn) {
                                    const uint32_t pos = pw + i, src = pos - off;
                                    uint32_t v;
                                    if ((int32_t) (src - (pw + base + 32 - kOutRing)) >= 0) v = ring[src & (kOutRing - 1)];
                                    else v = ld_far(out_al + src);
                                    ring[pos & (kOutRing - 1)] = (uint8_t) v;
                                }
                                __syncwarp();
                            }
                        }
                        else {
                            // periodic pattern: every byte of the piece repeats one of the `off` bytes in front of it
                            uint32_t m = (uint32_t) lane % off;
                            const uint32_t step = 32u % off;
                            for (uint32_t i = (uint32_t) lane; i < pn; i += 32) {
                                ring[(pw + i) & (kOutRing - 1)] = ring[(pw - off + m) & (kOutRing - 1)];
                                m += step;
                                if (m >= off) m -= off;
                            }
                            __syncwarp();
                        }
                        if (pw + pn - flushed >= (uint32_t) kFlushBytes) flush_to(ring, out_al, flushed, (pw + pn) & ~15u, lane);
                    }
                }
                if (endw - flushed >= (uint32_t) kFlushBytes) flush_to(ring, out_al, flushed, endw & ~15u, lane);
            }
            else if (r.w == kRecBegin) {
                blk = r.x;
                uint8_t *out = b.dst + b.dst_off[blk];
                const uint32_t oh = (uint32_t) ((uintptr_t) out & 15);
                out_al = out - oh;
                flushed = oh;
            }
            else if (r.w == kRecFallback) {
                const BlockDesc d = load_desc(b, blk);
                if (r.x == kFallbackWhole) Codec::general_whole(d, b, blk, lane);
                else {
                    // everything in front of the restart point must be in global memory (literal pieces of the sequence
                    // the general path decodes again may already be there: it rewrites the same bytes)
                    const uint32_t e = r.y + (uint32_t) ((uintptr_t) d.out & 15);
                    if ((int32_t) (e - flushed) > 0) flush_to(ring, out_al, flushed, e, lane);
                    Codec::general_from(d, b, blk, r.x, r.y, lane);
                }
                __syncwarp();
            }
            else {   // kRecExit
                return;
            }
        }
        cons += n;
        __syncwarp();
        if (lane == 0) {
            if (have_q) st_relaxed(&S.cons_q, last_q);
            st_release(&S.cons, cons);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// kernel body: warps [0, kSlots) execute, warp kSlots parses (one lane per slot)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void init_slot(Slot &S)
{
    S.prod = 0; S.cons = 0; S.cons_q = 0; S.abort = 0;
    for (int i = 0; i < kNChunk; i++) mbar_init(&S.mbar[i]);
}

template <class Codec, int kSlots>
__device__ __forceinline__ void run_warp(const AccBatch &b, Slot *slots, const int warp, const int lane)
{
    if (warp == kSlots) {
        if (lane < kSlots) parse_lane<Codec>(b, slots[lane]);
    }
    else exec_warp<Codec>(b, slots[warp], lane);
}

}  // namespace lzs
