// lz4_decode_v1.cuh -- warp-per-block LZ4 block decoder, bit-exact with Lz4RawDecompressor.java:35-198 including every
// reject decision and error offset.  Three layers, fastest first; each layer only takes what it can prove valid and leaves
// everything else (and all error reporting) to the next one:
//   1. multi-sequence steps: up to three sequences without length-extension bytes per 32-byte window (kFast >= 2)
//   2. medium steps: one sequence with a single extension byte per length (kFast == 3)
//   3. the general path: a restatement of the Java loop, sequence by sequence, with warp-wide copies
// Used by lz4_decompress_kernel (lz4.cu).
#pragma once
#include "acc_device.cuh"

namespace lz4v1 {

constexpr int kMinMatch = 4;
constexpr int kLastLiterals = 5;

// ------------------------------------------------------------------------------------------------
// Decode: one warp per block; `lane` is the caller's lane id, results go to out_len_base[idx] / status_base[idx].
// ------------------------------------------------------------------------------------------------
// PosT is the type of the positions: int32_t for blocks below 2 GiB (the normal case: half the registers and
// instructions), int64_t otherwise (no fast paths).  Sums that can leave the int32 range are formed in int64.
template <int kFast, typename PosT>
__device__ __forceinline__ void lz4_decode_impl(const uint8_t *__restrict__ in_arg, const PosT in_len, uint8_t *out_arg, const PosT out_cap,
                                                int64_t *out_len_base, int32_t *status_base, const uint32_t idx, const int lane,
                                                const PosT ip0 = 0, const PosT op0 = 0)
{
#define LZ4_FAIL(off, reason) do { if (lane == 0) { out_len_base[idx] = (int64_t) (off); status_base[idx] = ACC_STATUS(ACC_E_MALFORMED, reason); } return; } while (0)
    constexpr bool small = sizeof(PosT) == 4;
    const PosT fast_output_limit = out_cap - 8;
    PosT ip = ip0, op = op0;   // (ip0, op0) != (0, 0): resume at a sequence boundary
    // base pointers made opaque, so that the compiler keeps these two 64-bit values in registers instead of re-adding
    // kernel parameters (and holding the parts) inside the loops
    const uint8_t *in = in_arg;
    uint8_t *out = out_arg;
    asm volatile("" : "+l"(in));
    asm volatile("" : "+l"(out));
    __builtin_assume(__isGlobal(in));    // laundering hides the address space: say it again, or the accesses become generic
    __builtin_assume(__isGlobal(out));

    if (in_len == 0) LZ4_FAIL(0, ACC_R_INPUT_EMPTY);
    if (out_cap == 0) {
        if (in_len == 1 && in[0] == 0) { if (lane == 0) { out_len_base[idx] = 0; status_base[idx] = 0; } return; }
        if (lane == 0) { out_len_base[idx] = 0; status_base[idx] = ACC_STATUS(ACC_E_DST_TOO_SMALL, ACC_R_LZ4_ZERO_CAPACITY); }
        return;
    }

    // ---- fast path ------------------------------------------------------------------------------
    // Most sequences have no length-extension bytes (literal length < 15, match length < 19), i.e. they
    // produce at most 32 bytes.  For those, every lane resolves the source of ONE output byte directly
    // (a literal from the input, an older output byte, or -- when the match overlaps this sequence's own
    // literals / itself -- the literal it ultimately repeats) and the whole sequence is a single
    // load + store per lane.  The bounds below are exactly the conditions under which the Java decoder
    // takes its normal (non-final) path (Lz4RawDecompressor.java:82,168), so results are identical.
    while (ip < in_len) {
        if (small && kFast >= 2) {
            // ---- multi-sequence steps: up to three short sequences (no length-extension bytes) whose tokens, literals
            // and offsets all lie in the 32-byte input window at ip and whose output fits the 32 lanes.
            // Every lane first treats ITS byte as a token and works out what a sequence starting there would be
            // (literal length, output bytes, offset); the warp then follows the chain 0 -> next token -> next token
            // with three shuffles per sequence, and every lane resolves the source of one output byte: a literal of the
            // window, older output (one global load), or a byte another lane produces in this same step (taken by
            // shuffle once that lane has it; the dependency always points to a lower lane, so the loop ends after at
            // most three rounds, usually zero).
            // Bounds: every sequence ends <= ip + 32 <= in_len - 8 and <= op + 32 <= out_cap - 12, which are the Java
            // decoder's conditions for the normal (non-final) path (Lz4RawDecompressor.java:82,168); a sequence with a bad
            // offset is left for the next step to report (it then is sequence 0), so error offsets are unchanged.
            uint32_t ipw = (uint32_t) ip, opw = (uint32_t) op;
            while ((int32_t) (ipw + 40) <= in_len && (int32_t) (opw + 44) <= out_cap) {   // positions < 2^31 - 256: no wrap
                const uint32_t vb = __ldg(in + (ipw + (uint32_t) lane));
                const uint32_t tok = __shfl_sync(kFull, vb, 0);
                const uint32_t ll0 = tok >> 4, ml0 = tok & 15;
                if (ll0 == 15 || ml0 == 15) {
                    if (kFast != 3) break;
                    // ---- medium step: ONE sequence whose literal and/or match length carries a single extension byte
                    // (lengths up to 269 / 273; 92 % of the sequences the steps below cannot take).  Anything unusual --
                    // a second extension byte, the end-of-block rules, a bad offset, a match overlapping itself at a
                    // distance below 32 -- leaves through `break` BEFORE ip/op move, and the general path below decodes
                    // the sequence again from its token (and reports the error, if there is one).
                    uint32_t ll = ll0, ml = ml0, pos = 1;
                    if (ll == 15) {
                        const uint32_t x = __shfl_sync(kFull, vb, 1);
                        if (x == 255) break;
                        ll += x;
                        pos = 2;
                    }
                    const uint32_t lit_in = ipw + pos, lit_end = lit_in + ll;
                    // Java's normal-path conditions (Lz4RawDecompressor.java:82): literals end 8 bytes before the input end
                    // and 12 bytes before the output end
                    if (lit_end + 8 > (uint32_t) in_len || opw + ll + 12 > (uint32_t) out_cap) break;
                    const uint32_t wend = pos + ll;                              // window position of the offset bytes
                    uint32_t off, used = wend + 2, x = 0;
                    if (wend + 2 < 32) {                                         // offset and extension byte are in the window
                        off = __shfl_sync(kFull, vb, wend) | (__shfl_sync(kFull, vb, wend + 1) << 8);
                        x = __shfl_sync(kFull, vb, wend + 2);
                    }
                    else {
                        off = (uint32_t) __ldg(in + lit_end) | ((uint32_t) __ldg(in + (lit_end + 1)) << 8);
                        if (ml == 15) x = __ldg(in + (lit_end + 2));
                    }
                    if (ml == 15) {
                        if (x == 255) break;
                        ml += x;
                        used++;
                    }
                    ml += kMinMatch;
                    const uint32_t mop = opw + ll;                               // first output byte of the match
                    if (off - 1 >= mop || (off < 32 && off < ml) || mop + ml + 12 > (uint32_t) out_cap) break;
                    if (wend <= 32) {                                            // all literals are in the window
                        const uint32_t v = __shfl_sync(kFull, vb, lane + pos);
                        if ((uint32_t) lane < ll) out[opw + (uint32_t) lane] = (uint8_t) v;
                    }
                    else {
#pragma unroll 1
                        for (uint32_t i = lane; i < ll; i += 32) out[opw + i] = __ldg(in + (lit_in + i));
                    }
                    __syncwarp();
                    // 32 bytes per round; a round only reads bytes written at least 32 positions earlier, or (off >= ml,
                    // single round) bytes in front of the match
#pragma unroll 1
                    for (uint32_t base = 0; base < ml; base += 32) {
                        const uint32_t i = base + lane;
                        if (i < ml) out[mop + i] = out[mop + i - off];
                        __syncwarp();
                    }
                    ipw += used;
                    opw = mop + ml;
                    continue;
                }
                // per-lane candidates: what a sequence starting at this lane's byte would be
                const uint32_t ll_l = vb >> 4, ml_l = vb & 15;
                const uint32_t o_lo = __shfl_sync(kFull, vb, lane + 1 + ll_l);
                const uint32_t o_hi = __shfl_sync(kFull, vb, lane + 2 + ll_l);
                const uint32_t off_l = o_hi * 256 + o_lo;
                // output bytes of a sequence starting at this lane; >= 64 (never fits) when it needs the other paths
                uint32_t n_l = ll_l + ml_l + kMinMatch;
                if (vb >= 0xF0 || ml_l == 15 || (uint32_t) lane + ll_l > 29) n_l |= 64u;
                const uint32_t e0 = ll0 + ml0 + kMinMatch;                      // <= 32
                const uint32_t off0 = __shfl_sync(kFull, off_l, 0);
                if (off0 - 1 >= opw + ll0) { LZ4_FAIL((int64_t) ipw + ll0 + 3, ACC_R_OFFSET_OUTSIDE); }   // offset == 0 || offset > op
                const uint32_t nx1 = 3 + ll0;                                   // <= 17: always inside the window
                const uint32_t ll1 = __shfl_sync(kFull, ll_l, nx1), off1 = __shfl_sync(kFull, off_l, nx1);
                const uint32_t e1 = e0 + __shfl_sync(kFull, n_l, nx1);
                const bool v1 = e1 <= 32 && off1 - 1 < opw + e0 + ll1;
                const uint32_t nx2 = nx1 + 3 + ll1;                             // <= 32 when v1
                const uint32_t ll2 = __shfl_sync(kFull, ll_l, nx2), off2 = __shfl_sync(kFull, off_l, nx2);
                const uint32_t e2 = e1 + __shfl_sync(kFull, n_l, nx2);
                const bool v2 = v1 && nx2 < 32 && e2 <= 32 && off2 - 1 < opw + e1 + ll2;
                const uint32_t e = v2 ? e2 : v1 ? e1 : e0;                      // output bytes of this step
                const uint32_t nx = v2 ? nx2 + 3 + ll2 : v1 ? nx2 : nx1;        // input bytes of this step
                // which sequence produces output byte `lane`
                const bool k2 = v2 && (uint32_t) lane >= e1, k1 = v1 && (uint32_t) lane >= e0;
                const uint32_t sk = k2 ? nx2 : k1 ? nx1 : 0u;                   // token position in the window
                const uint32_t bk = k2 ? e1 : k1 ? e0 : 0u;                     // first output byte of the sequence
                const uint32_t lk = k2 ? ll2 : k1 ? ll1 : ll0;
                const uint32_t fk = k2 ? off2 : k1 ? off1 : off0;
                const uint32_t t = (uint32_t) lane - bk;
                const bool is_lit = t < lk;
                uint32_t val = __shfl_sync(kFull, vb, sk + 1 + t);              // the literal, if it is one
                int32_t m = (int32_t) t - (int32_t) lk;
                if (!is_lit && (uint32_t) m >= fk) m -= (int32_t) (fk * (((uint32_t) m * kRcp16[fk]) >> 16));   // m mod offset (offset < 32 here)
                const int32_t srel = (int32_t) (bk + lk) - (int32_t) fk + m;    // source, relative to op
                uint32_t need = ((uint32_t) lane < e && !is_lit) ? 256u : 0u;
                if (need && srel < 0) { val = out[opw + (uint32_t) srel]; need = 0; }   // opw + srel >= 0 (offsets checked)
                while (__any_sync(kFull, need)) {
                    const uint32_t w = __shfl_sync(kFull, val | need, srel);
                    if (need && !(w & 256u)) { val = w; need = 0; }
                }
                if ((uint32_t) lane < e) out[opw + (uint32_t) lane] = (uint8_t) val;
                __syncwarp();
                ipw += nx;
                opw += e;
            }
            ip = (PosT) ipw;
            op = (PosT) opw;
        }
        if (kFast == 1 && small && (int64_t) ip + 32 <= in_len && (int64_t) op + 44 <= out_cap) {
            // one coalesced 32-byte load: lane l holds input byte ip + l (token, literals, offset all inside)
            const uint32_t ipw = (uint32_t) ip, opw = (uint32_t) op;
            const uint32_t vb = __ldg(in + ipw + lane);
            const uint32_t tk = __shfl_sync(kFull, vb, 0);
            const uint32_t fll = tk >> 4, fml = tk & 15;
            if (fll != 15 && fml != 15) {
                const uint32_t foff = __shfl_sync(kFull, vb, fll + 1) | (__shfl_sync(kFull, vb, fll + 2) << 8);
                if (foff == 0 || foff > opw + fll) LZ4_FAIL((int64_t) ipw + fll + 3, ACC_R_OFFSET_OUTSIDE);
                const uint32_t total = fll + fml + kMinMatch;
                // source of output byte `lane`: literal (input byte ip+1+lane), or match byte m = lane - ll taken from
                // position rel (relative to op) = ll - offset + (m mod offset): rel >= 0 -> one of this sequence's own
                // literals, rel < 0 -> older output
                int32_t rel = (int32_t) lane;
                if ((uint32_t) lane >= fll) {
                    uint32_t m = (uint32_t) lane - fll;
                    if (m >= foff) m -= foff * ((m * kRcp16[foff]) >> 16);
                    rel = (int32_t) fll - (int32_t) foff + (int32_t) m;
                }
                uint32_t v = __shfl_sync(kFull, vb, (rel + 1) & 31);
                if ((uint32_t) lane < total) {
                    if (rel < 0) v = out[(int64_t) opw + rel];
                    out[opw + lane] = (uint8_t) v;
                }
                __syncwarp();
                ip = ipw + fll + 3;
                op = opw + total;
                continue;
            }
        }
        const uint32_t token = in[ip++];
        uint32_t ll = token >> 4;
        if (ll == 15) {
            if (ip >= in_len) LZ4_FAIL(ip, ACC_R_NONE);
            uint32_t v;
            do {
                v = in[ip++];
                ll += v;  // 32-bit wrap like the Java int
            }
            while (v == 255 && ip < in_len - 15);
        }
        if ((int32_t) ll < 0) LZ4_FAIL(ip, ACC_R_NONE);

        const int64_t lit_end = (int64_t) ip + (int64_t) ll;
        const int64_t lit_out_limit = (int64_t) op + (int64_t) ll;
        if (lit_out_limit > (int64_t) fast_output_limit - kMinMatch || lit_end > (int64_t) in_len - (2 + 1 + kLastLiterals)) {
            if (lit_out_limit > out_cap) LZ4_FAIL(ip, ACC_R_LAST_LITERAL_OUTSIDE);
            if (lit_end != in_len) LZ4_FAIL(ip, ACC_R_ALL_INPUT_CONSUMED);
            warp_copy(out + op, in + ip, ll, lane);
            op += (PosT) ll;
            break;
        }
        warp_copy(out + op, in + ip, ll, lane);
        op = (PosT) lit_out_limit;
        ip = (PosT) lit_end;

        const uint32_t offset = ld_u16le(in + ip);
        ip += 2;
        if ((int64_t) offset > (int64_t) op || offset == 0) LZ4_FAIL(ip, ACC_R_OFFSET_OUTSIDE);

        uint32_t ml = token & 15;
        if (ml == 15) {
            uint32_t v;
            do {
                if (ip > in_len - kLastLiterals) LZ4_FAIL(ip, ACC_R_NONE);
                v = in[ip++];
                ml += v;
            }
            while (v == 255);
        }
        ml += kMinMatch;
        if ((int32_t) ml < 0) LZ4_FAIL(ip, ACC_R_NONE);

        const int64_t match_out_limit = (int64_t) op + (int64_t) ml;
        if (match_out_limit > (int64_t) fast_output_limit - kMinMatch) {
            if (match_out_limit > (int64_t) out_cap - kLastLiterals) LZ4_FAIL(ip, ACC_R_LAST5_LITERALS);
        }
        __syncwarp();
        warp_match_copy(out + op, offset, ml, lane);
        __syncwarp();
        op = (PosT) match_out_limit;
    }
    if (lane == 0) { out_len_base[idx] = (int64_t) op; status_base[idx] = 0; }
#undef LZ4_FAIL
}

// kFast selects the fast path: 1 = one sequence per step, 2 = up to three sequences per step,
// 3 = 2 + medium steps for sequences with one length-extension byte.
template <int kFast = 3>
__device__ __forceinline__ void lz4_decode_block(const uint8_t *__restrict__ in, int64_t in_len, uint8_t *out, int64_t out_cap,
                                                 int64_t *out_len_base, int32_t *status_base, uint32_t idx, int lane)
{
    if (in_len < 0x7fffff00LL && out_cap < 0x7fffff00LL)
        lz4_decode_impl<kFast, int32_t>(in, (int32_t) in_len, out, (int32_t) out_cap, out_len_base, status_base, idx, lane);
    else
        lz4_decode_impl<0, int64_t>(in, in_len, out, out_cap, out_len_base, status_base, idx, lane);
}


}  // namespace lz4v1
