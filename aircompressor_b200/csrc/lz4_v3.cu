// lz4_v3.cu -- LZ4 block decoder with the 64 KiB output window in shared memory (sm_100a).
//
// Same contract as lz4_decompress_kernel (bit-exact with Lz4RawDecompressor.java:35-198).  One 256-thread CTA
// per block, two CTAs per SM.  The kernel is an *optimistic* decoder: whenever a block shows anything outside
// the common valid shape (a malformed token, an output-capacity corner case of the Java decoder, more than
// 64 KiB of output, an input larger than 128 KiB) it discards its work and warp 0 re-decodes the block with
// the exact serial decoder of lz4_decode_v1.cuh, so accept/reject decisions and error offsets never change.
//
// Per 16 KiB chunk of compressed input (staged in shared memory with coalesced 16-byte loads):
//   1. speculative parallel parse -- thread t owns the 64-byte input segment t and walks the tokens from a
//      guessed entry position; the exit of segment t-1 is the true entry of segment t, so entries are repaired
//      and segments re-parsed until the chain is consistent (LZ4 token streams re-synchronise quickly: 2-4
//      rounds on Silesia);
//   2. block-wide prefix sums over (sequences, output bytes) per segment give every sequence its table slot
//      and output position; sequences are materialised in batches of 1792 into a shared-memory table;
//   3. literals: one lane per sequence copies its literals input -> window (runs > 32 bytes are copied by a
//      whole warp) and marks the bytes final in a 1-bit-per-byte bitmap;
//   4. matches: lane per sequence, in rounds.  A match executes as soon as the bitmap says its source range is
//      final; its own output is then marked final.  Matches longer than 64 bytes are executed by the whole
//      warp.  The dependency depth of a 64 KiB Silesia block is ~90 sequences (3,700 sequences per block), so
//      most lanes find work in every round;
//   5. the window is flushed to HBM with 16-byte stores.
#include "acc_device.cuh"
#include "lz4_decode_v1.cuh"

__device__ unsigned long long g_lz4v3_stats[16];

namespace {

constexpr int kT = 256;
constexpr int kWin = 65536;
constexpr int kChunk = 16384;
constexpr int kSlack = 256;
constexpr int kSeg = kChunk / kT;      // 64 input bytes per parse segment
constexpr int kTab = 1792;
constexpr int kPer = kTab / kT;        // table slots per thread (strided ownership)
constexpr int kLongLit = 32;
constexpr int kLongMatch = 64;
constexpr int kMaxIn = 1 << 17;

struct Smem {
    uint8_t win[kWin];
    uint8_t inb[kChunk + kSlack + 32];
    uint32_t fin[kWin / 32];
    uint16_t t_lit[kTab], t_ll[kTab], t_ml[kTab], t_off[kTab], t_out[kTab];
    uint32_t exitp[kT];
    uint32_t scan[kT + 1];
    uint16_t long_list[kT];
    int n_long;
    int fallback;
};

struct In {
    const uint8_t *g;     // block input in global memory
    const uint8_t *s;     // s[p - c0] = staged copy of g[p] for c0 <= p < staged_end
    uint32_t c0, staged;  // staged = staged_end - c0
    __device__ __forceinline__ uint32_t rd(uint32_t p) const
    {
        const uint32_t r = p - c0;
        return r < staged ? s[r] : __ldg(g + p);
    }
};

struct Seq { uint32_t lit_pos, ll, ml, off, next; };

// one sequence at `pos` (< in_len).  0 = normal, 1 = final literal-only sequence, 2 = not the common valid shape.
// Follows Lz4RawDecompressor.java:58-138 (length decoding, the "last literals" input rule :82, the match-length guard :126).
__device__ __forceinline__ int decode_seq(const In &in, uint32_t pos, uint32_t in_len, Seq &q)
{
    const uint32_t token = in.rd(pos);
    uint32_t p = pos + 1, ll = token >> 4;
    if (ll == 15) {
        if (p >= in_len) return 2;
        uint32_t v;
        do { v = in.rd(p++); ll += v; } while (v == 255 && p + 15 < in_len);
    }
    q.lit_pos = p;
    q.ll = ll;
    const uint32_t lit_end = p + ll;
    if (lit_end + 8 > in_len) {
        if (lit_end != in_len) return 2;
        q.ml = 0; q.off = 0; q.next = in_len;
        return 1;
    }
    q.off = in.rd(lit_end) | (in.rd(lit_end + 1) << 8);
    p = lit_end + 2;
    uint32_t ml = token & 15;
    if (ml == 15) {
        uint32_t v;
        do { if (p + 5 > in_len) return 2; v = in.rd(p++); ml += v; } while (v == 255);
    }
    q.ml = ml + 4;
    q.next = p;
    return 0;
}

__device__ __forceinline__ void set_final(uint32_t *fin, uint32_t a, uint32_t b)   // bits [a, b)
{
    if (a >= b) return;
    uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
    const uint32_t m0 = 0xFFFFFFFFu << (a & 31), m1 = 0xFFFFFFFFu >> (31 - ((b - 1) & 31));
    if (w0 == w1) { atomicOr(fin + w0, m0 & m1); return; }
    atomicOr(fin + w0, m0);
    for (uint32_t w = w0 + 1; w < w1; w++) fin[w] = 0xFFFFFFFFu;
    atomicOr(fin + w1, m1);
}

__device__ __forceinline__ bool is_final(const volatile uint32_t *fin, uint32_t a, uint32_t b)   // all bits of [a, b) set?
{
    if (a >= b) return true;
    uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
    const uint32_t m0 = 0xFFFFFFFFu << (a & 31), m1 = 0xFFFFFFFFu >> (31 - ((b - 1) & 31));
    if (w0 == w1) return (fin[w0] & m0 & m1) == (m0 & m1);
    if ((fin[w0] & m0) != m0) return false;
    for (uint32_t w = w0 + 1; w < w1; w++) if (fin[w] != 0xFFFFFFFFu) return false;
    return (fin[w1] & m1) == m1;
}

// exclusive scan of one value per thread; total in *total (two barriers)
__device__ __forceinline__ uint32_t block_scan(Smem &sm, uint32_t v, uint32_t *total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) sm.scan[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < kT / 32; w++) { uint32_t s = sm.scan[w]; if (w < warp) base += s; tot += s; }
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(kT, 2) lz4_decompress_v3_kernel(AccBatch b)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned int s_idx;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(b.work_counter, 1u);
        __syncthreads();
        const unsigned int idx = s_idx;
        if ((int64_t) idx >= b.n) break;
        const uint8_t *gin = b.src + b.src_off[idx];
        const int64_t in_len64 = b.src_len[idx];
        uint8_t *gout = b.dst + b.dst_off[idx];
        const int64_t out_cap64 = b.dst_cap[idx];

        bool fallback = !(in_len64 >= 1 && in_len64 <= kMaxIn && out_cap64 >= 1);
        uint32_t out_total = 0;
        long long t_parse = 0, t_fill = 0, t_lit = 0, t_match = 0, t_flush = 0, t0 = clock64(), tb = t0;
        unsigned n_prounds = 0, n_mrounds = 0, n_seq = 0;
#define TICK(acc) do { long long now_ = clock64(); acc += now_ - tb; tb = now_; } while (0)
        if (!fallback) {
            const uint32_t in_len = (uint32_t) in_len64;
            const uint32_t out_cap = out_cap64 > 0x7fffffff ? 0x7fffffffu : (uint32_t) out_cap64;
            const uint32_t out_lim = out_cap < kWin ? out_cap : kWin;
            for (int i = tid; i < kWin / 32; i += kT) sm.fin[i] = 0;
            if (tid == 0) { sm.fallback = 0; sm.n_long = 0; }
            uint32_t carry_entry = 0;
            bool stream_done = false;

            for (uint32_t c0 = 0; c0 < in_len && !stream_done && !fallback; c0 += kChunk) {
                // ---- stage the chunk: aligned 16-byte loads ----
                const uint32_t stage_end = min(c0 + kChunk + kSlack, in_len);
                const uintptr_t ga = (uintptr_t) (gin + c0);
                const uint32_t mis = (uint32_t) (ga & 15);
                {
                    const uint4 *g16 = reinterpret_cast<const uint4 *>(ga - mis);
                    uint4 *s16 = reinterpret_cast<uint4 *>(sm.inb);
                    const uint32_t nvec = (stage_end - c0 + mis + 15) >> 4;
                    for (uint32_t v = tid; v < nvec; v += kT) s16[v] = __ldg(g16 + v);
                }
                In in;
                in.g = gin; in.s = sm.inb + mis; in.c0 = c0; in.staged = stage_end - c0;
                __syncthreads();
                if (carry_entry >= c0 + kChunk) continue;   // a long literal run covers this whole chunk (uniform)
                TICK(t_fill);

                // ---- 1. speculative parse ----
                const uint32_t seg_start = min(c0 + tid * kSeg, in_len);
                const uint32_t seg_end = min(seg_start + kSeg, in_len);
                uint32_t entry = tid == 0 ? carry_entry : seg_start;
                uint32_t exitp = 0, cnt = 0, ob = 0;
                bool bad = false, has_final = false, need = true;
                for (;;) {
                    if (need) {
                        uint32_t pos = entry;
                        cnt = 0; ob = 0; bad = false; has_final = false;
                        while (pos < seg_end) {
                            Seq q;
                            const int r = decode_seq(in, pos, in_len, q);
                            if (r == 2) { bad = true; pos = in_len; break; }
                            cnt++;
                            ob += q.ll + q.ml;
                            pos = q.next;
                            if (r == 1) { has_final = true; break; }
                        }
                        exitp = pos;
                    }
                    sm.exitp[tid] = exitp;
                    __syncthreads();
                    const uint32_t true_entry = tid == 0 ? carry_entry : sm.exitp[tid - 1];
                    need = true_entry != entry;
                    entry = true_entry;
                    n_prounds++;
                    if (!__syncthreads_or(need)) break;
                }
                TICK(t_parse);
                carry_entry = sm.exitp[kT - 1];
                if (__syncthreads_or(bad)) { fallback = true; break; }
                stream_done = __syncthreads_or(has_final) != 0;

                // ---- 2. slots and output positions ----
                uint32_t n_chunk, chunk_out;
                const uint32_t seq_base = block_scan(sm, cnt, &n_chunk);
                uint32_t out_pos = out_total + block_scan(sm, ob, &chunk_out);
                if (out_total + chunk_out > out_lim) { fallback = true; break; }
                uint32_t pos = entry, next_idx = seq_base, remaining = cnt;
                bool my_bad = false;

                for (uint32_t b0 = 0; b0 < n_chunk; b0 += kTab) {
                    const uint32_t bn = min((uint32_t) kTab, n_chunk - b0);
                    // fill this batch of the table
                    while (remaining > 0 && next_idx < b0 + kTab) {
                        Seq q;
                        const int r = decode_seq(in, pos, in_len, q);
                        const uint32_t slot = next_idx - b0;
                        if (q.ll > 65535 || q.ml > 65535) my_bad = true;
                        if (r == 0) {
                            // output-side rules of the Java decoder for a non-final sequence (:82, :168-171) and the offset check (:116-119)
                            const uint32_t ms = out_pos + q.ll;
                            if (ms + 12 > out_cap || ms + q.ml + 5 > out_cap || q.off == 0 || q.off > ms) my_bad = true;
                        }
                        sm.t_lit[slot] = (uint16_t) (q.lit_pos - c0);
                        sm.t_ll[slot] = (uint16_t) q.ll;
                        sm.t_ml[slot] = (uint16_t) q.ml;
                        sm.t_off[slot] = (uint16_t) q.off;
                        sm.t_out[slot] = (uint16_t) out_pos;
                        out_pos += q.ll + q.ml;
                        pos = q.next;
                        next_idx++;
                        remaining--;
                    }
                    if (__syncthreads_or(my_bad)) { fallback = true; break; }
                    TICK(t_fill);
                    n_seq += bn;

                    // ---- 3. literals ----
                    for (uint32_t i = tid; i < bn; i += kT) {
                        const uint32_t ll = sm.t_ll[i], o = sm.t_out[i], lp = c0 + sm.t_lit[i];
                        if (ll <= kLongLit) {
                            for (uint32_t k = 0; k < ll; k++) sm.win[o + k] = (uint8_t) in.rd(lp + k);
                            set_final(sm.fin, o, o + ll);
                        }
                        else {
                            const int li = atomicAdd(&sm.n_long, 1);
                            if (li < kT) sm.long_list[li] = (uint16_t) i;
                            else {   // list full: copy it here
                                for (uint32_t k = 0; k < ll; k++) sm.win[o + k] = (uint8_t) in.rd(lp + k);
                                set_final(sm.fin, o, o + ll);
                            }
                        }
                    }
                    __syncthreads();
                    {
                        const int nl = min(sm.n_long, kT);
                        for (int li = warp; li < nl; li += kT / 32) {
                            const uint32_t i = sm.long_list[li];
                            const uint32_t ll = sm.t_ll[i], o = sm.t_out[i], lp = c0 + sm.t_lit[i];
                            warp_copy(sm.win + o, gin + lp, ll, lane);
                            if (lane == 0) set_final(sm.fin, o, o + ll);
                        }
                    }
                    __syncthreads();
                    if (tid == 0) sm.n_long = 0;
                    TICK(t_lit);

                    // ---- 4. matches, in rounds ----
                    int cur = 0;
                    const int nmine = (int) ((bn > (uint32_t) tid) ? (bn - tid + kT - 1) / kT : 0);
                    for (;;) {
                        uint32_t ms = 0, off = 0, ml = 0;
                        bool want_long = false;
                        while (cur < nmine) {
                            const uint32_t i = tid + cur * kT;
                            ml = sm.t_ml[i];
                            if (ml == 0) { cur++; continue; }
                            ms = (uint32_t) sm.t_out[i] + sm.t_ll[i];
                            off = sm.t_off[i];
                            const uint32_t src = ms - off;
                            if (!is_final(sm.fin, src, min(src + ml, ms))) break;
                            if (ml > kLongMatch) { want_long = true; break; }
                            for (uint32_t k = 0; k < ml; k++) sm.win[ms + k] = sm.win[src + k];
                            __threadfence_block();
                            set_final(sm.fin, ms, ms + ml);
                            cur++;
                        }
                        // long matches: the whole warp copies, the owner publishes
                        unsigned lm = __ballot_sync(kFull, want_long);
                        while (lm) {
                            const int l = __ffs(lm) - 1;
                            lm &= lm - 1;
                            const uint32_t ms_ = __shfl_sync(kFull, ms, l), off_ = __shfl_sync(kFull, off, l), ml_ = __shfl_sync(kFull, ml, l);
                            uint8_t *d = sm.win + ms_;
                            const uint8_t *s = d - off_;
                            if (off_ >= 32) {
                                for (uint32_t base = 0; base < ml_; base += 32) {
                                    const uint32_t k = base + lane;
                                    if (k < ml_) d[k] = s[k];
                                    __syncwarp();
                                }
                            }
                            else {
                                uint32_t m = lane % off_;
                                const uint32_t step = 32 % off_;
                                for (uint32_t k = lane; k < ml_; k += 32) {
                                    d[k] = s[m];
                                    m += step;
                                    if (m >= off_) m -= off_;
                                }
                            }
                            __syncwarp();
                            if (lane == l) { __threadfence_block(); set_final(sm.fin, ms_, ms_ + ml_); cur++; }
                        }
                        n_mrounds++;
                        if (!__syncthreads_or(cur < nmine)) break;
                    }
                    TICK(t_match);
                }
                if (fallback) break;
                out_total += chunk_out;
            }
            if (!fallback && !stream_done) fallback = true;   // no final literal sequence: let the exact decoder report it
        }

        __syncthreads();
        if (!fallback) {
            // ---- 5. flush ----
            const uint32_t per = ((out_total + kT / 32 - 1) / (kT / 32) + 15) & ~15u;
            const uint32_t a = min(per * warp, out_total), e = min(a + per, out_total);
            if (a < e) warp_copy(gout + a, sm.win + a, e - a, lane);
            if (tid == 0) { b.out_len[idx] = out_total; b.status[idx] = 0; }
            TICK(t_flush);
        }
        else if (warp == 0) {
            lz4v1::lz4_decode_block(gin, in_len64, gout, out_cap64, b.out_len, b.status, (uint32_t) idx, lane);
        }
        if (tid == 0) {
            atomicAdd(&g_lz4v3_stats[0], 1ull); atomicAdd(&g_lz4v3_stats[1], fallback ? 1ull : 0ull);
            atomicAdd(&g_lz4v3_stats[2], (unsigned long long) t_parse); atomicAdd(&g_lz4v3_stats[3], (unsigned long long) t_fill);
            atomicAdd(&g_lz4v3_stats[4], (unsigned long long) t_lit); atomicAdd(&g_lz4v3_stats[5], (unsigned long long) t_match);
            atomicAdd(&g_lz4v3_stats[6], (unsigned long long) t_flush); atomicAdd(&g_lz4v3_stats[7], (unsigned long long) n_prounds);
            atomicAdd(&g_lz4v3_stats[8], (unsigned long long) n_mrounds); atomicAdd(&g_lz4v3_stats[9], (unsigned long long) (clock64() - t0));
            atomicAdd(&g_lz4v3_stats[10], (unsigned long long) n_seq);
        }
#undef TICK
    }
}

}  // namespace

extern "C" void acc_debug_lz4v3_stats(unsigned long long *out16)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_lz4v3_stats, sizeof(unsigned long long) * 16);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_lz4v3_stats, z, sizeof(z));
}

void acc_launch_lz4_decompress_v3(const AccBatch &b, int sm_count, cudaStream_t st)
{
    const int smem = (int) sizeof(Smem);
    cudaFuncSetAttribute(lz4_decompress_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int64_t ctas = b.n;
    const int64_t max_ctas = (int64_t) sm_count * 2;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    lz4_decompress_v3_kernel<<<(unsigned) ctas, kT, smem, st>>>(b);
}
