// snappy_decode.cuh -- device functions of the Snappy raw-format decoder: the element-by-element restatement of
// SnappyRawDecompressor.java:35-321 with the warp-step fast paths in front of it, and the Snappy parse side of the record
// path (lz_records.cuh).  Included by snappy.cu and by the host emulation (tests/host/lzs_emu.cpp).
#pragma once
#include "acc_device.cuh"
#include "lz_records.cuh"

namespace snappydec {

// SnappyRawDecompressor.java:238-271 opLookupTable as a formula: bits 0-7 length, 8-10 offset/256,
// 11-13 trailer bytes.
__device__ __forceinline__ uint32_t snappy_op_entry(uint32_t op)
{
    uint32_t kind = op & 3, hi = op >> 2;
    if (kind == 0) return hi < 60 ? hi + 1 : (((hi - 59) << 11) | 1);
    if (kind == 1) return (1u << 11) | ((op >> 5) << 8) | (4 + (hi & 7));
    if (kind == 2) return (2u << 11) | (hi + 1);
    return (4u << 11) | (hi + 1);
}

// Java readUncompressedLength (SnappyRawDecompressor.java:277-321). Returns status word (0 = ok).
__device__ __forceinline__ int32_t snappy_read_length(const uint8_t *in, int64_t in_len, uint32_t *result_out, int *bytes_read, int64_t *err_off)
{
    uint32_t result = 0;
    int n = 0;
    for (int shift = 0;; shift += 7) {
        if (n >= in_len) { *err_off = in_len - n; return ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_TRUNCATED); }
        uint32_t b = in[n++];
        result |= (b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        if (shift == 28) { *err_off = n; return ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_VARINT_HIGHBIT); }
    }
    if ((int32_t) result < 0) { *err_off = 0; return ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_NEG_LENGTH); }
    *result_out = result;
    *bytes_read = n;
    return 0;
}

// The element loop of SnappyRawDecompressor.uncompressAll (:70-220), resumable at an element boundary: `in` points behind
// the length preamble, (ip0, op0) is where decoding (re)starts -- (0, 0) for a whole block.
// kMulti: try multi-element steps (up to four elements per warp step) before the pair path; kPair: the pair path.
template <bool kMulti, bool kPair = true>
__device__ __forceinline__ void snappy_decode_from(const uint8_t *__restrict__ in, const int64_t in_len, uint8_t *out, const int64_t out_cap, const uint32_t expected,
                                                   const int64_t ip0, const int64_t op0, int64_t *out_len, int32_t *status, int lane)
{
#define SN_FAIL(off) do { if (lane == 0) { *out_len = (off); *status = ACC_STATUS(ACC_E_MALFORMED, ACC_R_NONE); } return; } while (0)
    // base pointers made opaque so the compiler keeps the two 64-bit sums in registers (see lz4_decode_v1.cuh)
    asm volatile("" : "+l"(in));
    asm volatile("" : "+l"(out));
    __builtin_assume(__isGlobal(in));
    __builtin_assume(__isGlobal(out));
    const int64_t fast_output_limit = out_cap - 8;
    int64_t ip = ip0, op = op0;

    const bool small = in_len < 0x7fffff00LL && out_cap < 0x7fffff00LL;
    // multi-element steps run while ip <= ip_lim && op <= op_lim (-1: never)
    const int32_t ip_lim = (kMulti && small && in_len >= 32 && out_cap >= 32) ? (int32_t) in_len - 32 : -1;
    const int32_t op_lim = (kMulti && small && in_len >= 32 && out_cap >= 32) ? (int32_t) out_cap - 32 : -1;
    while (ip < in_len) {
        if (kMulti) {
            uint32_t ipw = (uint32_t) ip, opw = (uint32_t) op;
            while ((int32_t) ipw <= ip_lim && (int32_t) opw <= op_lim) {
                const uint32_t vb = __ldg(in + (ipw + (uint32_t) lane));
                // ---- multi-element step: up to four elements (literals of < 32 bytes, 1- and 2-byte-offset copies) that
                // lie completely in the 32-byte window and produce at most 32 bytes together.  Every lane first decodes
                // ITS byte as if it were a tag (output bytes, input bytes, offset), the warp follows the chain
                // tag -> next tag with two shuffles per element, and every lane resolves the source of one output byte:
                // a literal of the window, older output (one global load), or a byte another lane produces in this step
                // (taken by shuffle once that lane has it; dependencies point to lower lanes).  An element is only taken
                // when it is valid under SnappyRawDecompressor.java:89-216 (offset != 0, offset <= op, output fits);
                // anything else ends the chain and is left to the paths below, which report errors at the same offsets.
                const uint32_t kind = vb & 3, hi = vb >> 2;
                const uint32_t b1 = __shfl_sync(kFull, vb, lane + 1), b2 = __shfl_sync(kFull, vb, lane + 2);
                uint32_t outn = hi + 1, adv = 3, off = b1 | (b2 << 8);                  // 2-byte-offset copy
                if (kind == 1) { outn = 4 + (hi & 7); adv = 2; off = ((vb >> 5) << 8) | b1; }
                if (kind == 0) { adv = hi + 2; off = 0; }                               // literal: offset 0 marks it
                const bool usable = kind != 3 && !(kind == 0 && hi >= 60) && !(kind != 0 && off == 0) && (uint32_t) lane + adv <= 32 && outn <= 32;
                const uint32_t A = (usable ? outn : 127u) | ((adv & 63) << 8);
                const uint32_t a0 = __shfl_sync(kFull, A, 0), o0 = __shfl_sync(kFull, off, 0);
                const uint32_t n0 = a0 & 127, x1 = a0 >> 8;
                const bool ok0 = n0 <= 32 && o0 <= opw;                                // copies: 1 <= offset <= op (0 = literal)
                const uint32_t a1 = __shfl_sync(kFull, A, x1), o1 = __shfl_sync(kFull, off, x1);
                const uint32_t e1 = n0 + (a1 & 127), x2 = x1 + (a1 >> 8);
                const bool v1 = ok0 && x1 < 32 && e1 <= 32 && o1 <= opw + n0;
                if (!v1) break;
                {
                    const uint32_t a2 = __shfl_sync(kFull, A, x2), o2 = __shfl_sync(kFull, off, x2);
                    const uint32_t e2 = e1 + (a2 & 127), x3 = x2 + (a2 >> 8);
                    const bool v2 = x2 < 32 && e2 <= 32 && o2 <= opw + e1;
                    const uint32_t a3 = __shfl_sync(kFull, A, x3), o3 = __shfl_sync(kFull, off, x3);
                    const uint32_t e3 = e2 + (a3 & 127), x4 = x3 + (a3 >> 8);
                    const bool v3 = v2 && x3 < 32 && e3 <= 32 && o3 <= opw + e2;
                    const uint32_t e = v3 ? e3 : v2 ? e2 : e1;                          // output bytes of this step
                    const uint32_t nx = v3 ? x4 : v2 ? x3 : x2;                         // input bytes of this step
                    // which element produces output byte `lane`
                    const bool k3 = v3 && (uint32_t) lane >= e2, k2 = v2 && (uint32_t) lane >= e1, k1 = (uint32_t) lane >= n0;
                    const uint32_t sk = k3 ? x3 : k2 ? x2 : k1 ? x1 : 0u;               // tag position in the window
                    const uint32_t bk = k3 ? e2 : k2 ? e1 : k1 ? n0 : 0u;               // first output byte of the element
                    const uint32_t fk = k3 ? o3 : k2 ? o2 : k1 ? o1 : o0;               // offset (0: literal)
                    const uint32_t t = (uint32_t) lane - bk;
                    uint32_t val = __shfl_sync(kFull, vb, sk + 1 + t);                  // the literal byte, if it is one
                    uint32_t m = t;
                    if (fk != 0 && m >= fk) m -= fk * ((m * kRcp16[fk]) >> 16);         // m mod offset (offset < 32 here)
                    const int32_t srel = (int32_t) (bk + m) - (int32_t) fk;             // source, relative to op
                    uint32_t need = ((uint32_t) lane < e && fk != 0) ? 256u : 0u;
                    if (need && srel < 0) { val = out[opw + (uint32_t) srel]; need = 0; }   // opw + srel >= 0 (offsets checked)
                    while (__any_sync(kFull, need)) {
                        const uint32_t w = __shfl_sync(kFull, val | need, srel);
                        if (need && !(w & 256u)) { val = w; need = 0; }
                    }
                    if ((uint32_t) lane < e) out[opw + lane] = (uint8_t) val;
                    __syncwarp();
                    ipw += nx;
                    opw += e;
                }
            }
            ip = ipw;
            op = opw;
            if (ip >= in_len) break;
        }
        // ---- fast path: [literal of <= 27 bytes] + [one 1- or 2-byte-offset copy], parsed from one coalesced 32-byte load.
        // Every output byte is resolved independently (a literal byte of this step, or older output through the periodic
        // source formula), so the step is one load and one store per lane and 32-byte chunk.  The bounds make the elements
        // valid under SnappyRawDecompressor.java:89-216; anything else goes to the element-by-element path below.
        if (kPair && small && ip + 32 <= in_len) {
            const uint32_t ipw = (uint32_t) ip, opw = (uint32_t) op;
            const uint32_t vb = __ldg(in + ipw + lane);
            const uint32_t t0 = __shfl_sync(kFull, vb, 0);
            uint32_t L = 0, p = 0;
            bool ok = true;
            if ((t0 & 3) == 0) {
                const uint32_t n = t0 >> 2;
                if (n <= 26) { L = n + 1; p = 1 + L; } else ok = false;
            }
            if (ok) {
                const uint32_t tag = __shfl_sync(kFull, vb, p);
                const uint32_t b1 = __shfl_sync(kFull, vb, (p + 1) & 31), b2 = __shfl_sync(kFull, vb, (p + 2) & 31);
                const uint32_t kind = tag & 3;
                uint32_t clen = 0, coff = 1, adv = p;
                if (kind == 1) { clen = 4 + ((tag >> 2) & 7); coff = ((tag >> 5) << 8) | b1; adv = p + 2; }
                else if (kind == 2) { clen = (tag >> 2) + 1; coff = b1 | (b2 << 8); adv = p + 3; }
                else if (L == 0) ok = false;          // long literal / 4-byte-offset copy first: slow path
                const uint32_t total = L + clen;
                if (ok && coff != 0 && coff <= opw + L && (uint64_t) opw + total <= (uint64_t) out_cap) {
                    for (uint32_t c = 0; c < total; c += 32) {
                        const uint32_t j = c + lane;
                        int32_t rel = (int32_t) j;     // position relative to op of the byte to copy (literal: itself, from the window)
                        if (j >= L) {
                            uint32_t m = j - L;
                            if (m >= coff) m = m % coff;
                            rel = (int32_t) L - (int32_t) coff + (int32_t) m;
                        }
                        uint32_t v = __shfl_sync(kFull, vb, (rel + 1) & 31);
                        if (j < total) {
                            if (rel < 0) v = out[(int64_t) opw + rel];
                            out[opw + j] = (uint8_t) v;
                        }
                    }
                    __syncwarp();
                    ip = ipw + adv;
                    op = opw + total;
                    continue;
                }
            }
        }
        const uint32_t opc = in[ip++];
        const uint32_t entry = snappy_op_entry(opc);
        const int trailer_bytes = (int) (entry >> 11);
        if (!(ip + 4 < in_len)) {
            if (ip + trailer_bytes > in_len) SN_FAIL(ip);
        }
        uint32_t trailer = 0;
        for (int i = trailer_bytes - 1; i >= 0; i--) trailer = (trailer << 8) | in[ip + i];
        if ((int32_t) trailer < 0) SN_FAIL(ip);
        ip += trailer_bytes;
        const uint32_t length = entry & 0xff;

        if ((opc & 3) == 0) {
            const uint32_t ll = length + trailer;
            if ((int32_t) ll < 0) SN_FAIL(ip);
            const int64_t lit_out_limit = op + (int64_t) ll;
            if (lit_out_limit > fast_output_limit || ip + (int64_t) ll > in_len - 8) {
                if (lit_out_limit > out_cap || ip + (int64_t) ll > in_len) SN_FAIL(ip);
            }
            warp_copy(out + op, in + ip, ll, lane);
            __syncwarp();   // later steps read these bytes through other lanes
            ip += ll;
            op = lit_out_limit;
        }
        else {
            const uint32_t moff = (entry & 0x700) + trailer;
            if ((int32_t) moff <= 0) SN_FAIL(ip);
            if ((int64_t) moff > op || op + (int64_t) length > out_cap) SN_FAIL(ip);
            __syncwarp();
            warp_match_copy(out + op, moff, length, lane);
            __syncwarp();
            op += length;
        }
    }
    if ((int64_t) expected != op) {
        if (lane == 0) { *out_len = 0; *status = ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_LEN_MISMATCH); }
        return;
    }
    if (lane == 0) { *out_len = expected; *status = 0; }
#undef SN_FAIL
}

// preamble (SnappyRawDecompressor.java:35-68) + element loop
template <bool kMulti, bool kPair = true>
__device__ __forceinline__ void snappy_decode_block(const uint8_t *__restrict__ in0, int64_t in_len0, uint8_t *out, int64_t out_cap,
                                                    int64_t *out_len, int32_t *status, int lane, int64_t ip0 = 0, int64_t op0 = 0)
{
    uint32_t expected = 0;
    int br = 0;
    int64_t eoff = 0;
    int32_t st = snappy_read_length(in0, in_len0, &expected, &br, &eoff);
    if (st != 0) { if (lane == 0) { *out_len = eoff; *status = st; } return; }
    if ((int64_t) expected > out_cap) {
        if (lane == 0) { *out_len = 0; *status = ACC_STATUS(ACC_E_DST_TOO_SMALL, ACC_R_SNAPPY_LEN_GT_CAP); }
        return;
    }
    snappy_decode_from<kMulti, kPair>(in0 + br, in_len0 - br, out, out_cap, expected, ip0, op0, out_len, status, lane);
}

// ------------------------------------------------------------------------------------------------
// Record path (lz_records.cuh): the Snappy side of the parse lane.  A record is a literal, a copy, or a literal of
// <= 60 bytes together with the 1- or 2-byte-offset copy that follows it.  An element is taken only when it is valid on
// the Java decoder's fast path (SnappyRawDecompressor.java:89-216: literal ends >= 8 bytes before both limits;
// 1 <= offset <= output position; copy output fits) and all its bytes lie >= 16 bytes before the input end; anything else
// returns kFallback at the element, where snappy_decode_from resumes.  The length preamble (:35-68) is parsed first; a
// block whose preamble is not plainly valid goes to the general path whole.
// ------------------------------------------------------------------------------------------------
struct SnappyRecords {
    struct Parse {
        int32_t ip, op;            // next element, next output byte (positions behind the preamble)
        int32_t el_ip, el_op;      // where the step decoder takes over: the element that was not recorded (el_ip == -1: the whole block)
        bool started;              // the preamble has been read
    };
    static __device__ __forceinline__ void begin(Parse &P) { P.ip = 0; P.op = 0; P.el_ip = -1; P.el_op = 0; P.started = false; }
    static __device__ __forceinline__ uint32_t resume_ip(const Parse &P) { return (uint32_t) P.el_ip; }
    static __device__ __forceinline__ uint32_t resume_op(const Parse &P) { return (uint32_t) P.el_op; }

    // One element (or a literal of <= 60 bytes with the 1- / 2-byte-offset copy behind it): all checks, then its record(s).
    // Nothing is recorded before the element is known to be valid, so a hand-over always happens at an element.
    static __device__ __forceinline__ int parse_one(Parse &P, lzs::ParseCtx &C, const int row)
    {
        if (!P.started) {
            // varint preamble: in_len >= 32 here, so its <= 5 bytes exist
            uint32_t result = 0;
            int n = 0;
            C.ensure(0);
            for (int shift = 0;; shift += 7) {
                const uint32_t v = C.byte(n++);
                result |= (v & 0x7f) << shift;
                if (!(v & 0x80)) break;
                if (shift == 28) return lzs::kFallback;
            }
            if ((int32_t) result < 0 || (int32_t) result > C.out_cap) return lzs::kFallback;
            C.in += n;                 // positions from here on are relative to the first element
            C.in_len -= n;
            C.head = (uint32_t) ((uintptr_t) C.in & 31);
            C.win_chunks = (C.head + (uint32_t) C.in_len + 31) >> 5;
            C.win_tag = ~0u;
            P.started = true;
        }
        const int32_t safe_end = C.in_len - 16;
        const int32_t ip = P.ip;
        P.el_ip = ip; P.el_op = P.op;
        if (C.n_rec + 2 > row) return lzs::kRowFull;
        if (ip >= safe_end) return lzs::kFallback;
        C.ensure(ip);                                    // the tag and its <= 4 trailer bytes
        const uint32_t tag = C.byte(ip);
        if ((tag & 3) == 0) {
            const uint32_t hi = tag >> 2;
            int32_t p = ip + 1;
            uint32_t ll = hi + 1;
            if (hi >= 60) {
                const int nb = (int) hi - 59;
                uint32_t v = 0;
                for (int i = 0; i < nb; i++) v |= C.byte(p + i) << (8 * i);
                p += nb;
                if (v >= (1u << 24)) return lzs::kFallback;
                ll = v + 1;
            }
            if (p + (int32_t) ll + 8 > C.in_len || P.op + (int32_t) ll + 8 > C.out_cap) return lzs::kFallback;
            if (hi >= 60) {
                // a long literal travels as literal-only records (12-bit length field)
                int32_t lp = p, rem = (int32_t) ll;
                while (rem > 0) {
                    if (C.n_rec + 2 > row) return lzs::kRowFull;       // the pieces already recorded are written again by the step decoder: same bytes
                    const int32_t n = rem < lzs::kMaxLitPiece ? rem : lzs::kMaxLitPiece;
                    if (!C.emit(lp, (uint32_t) n, 0, lzs::kNoOffset)) return lzs::kFallback;
                    lp += n; rem -= n;
                }
                P.ip = p + (int32_t) ll;
                P.op += (int32_t) ll;
                return lzs::kMore;
            }
            // short literal: take the copy behind it into the same record when there is one
            const int32_t q = p + (int32_t) ll;
            if (q < safe_end) {
                C.ensure(q);                             // the literal's bytes are skipped, not read: the window moves on
                const uint32_t t2 = C.byte(q);
                const uint32_t k2 = t2 & 3;
                if (k2 == 1 || k2 == 2) {
                    const uint32_t b1 = C.byte(q + 1);
                    uint32_t ml, off;
                    int32_t adv;
                    if (k2 == 1) { ml = 4 + ((t2 >> 2) & 7); off = ((t2 >> 5) << 8) | b1; adv = 2; }
                    else { ml = (t2 >> 2) + 1; off = b1 | (C.byte(q + 2) << 8); adv = 3; }
                    const int32_t mop = P.op + (int32_t) ll;
                    if (off != 0 && (int32_t) off <= mop && mop + (int32_t) ml + 8 <= C.out_cap) {
                        if (!C.emit(p, ll, ml, off)) return lzs::kFallback;
                        P.ip = q + adv;
                        P.op = mop + (int32_t) ml;
                        return lzs::kMore;
                    }
                }
            }
            if (!C.emit(p, ll, 0, lzs::kNoOffset)) return lzs::kFallback;
            P.ip = q;
            P.op += (int32_t) ll;
            return lzs::kMore;
        }
        const uint32_t kind = tag & 3;
        const uint32_t b1 = C.byte(ip + 1);
        uint32_t ml, off;
        int32_t adv;
        if (kind == 1) { ml = 4 + ((tag >> 2) & 7); off = ((tag >> 5) << 8) | b1; adv = 2; }
        else if (kind == 2) { ml = (tag >> 2) + 1; off = b1 | (C.byte(ip + 2) << 8); adv = 3; }
        else {
            ml = (tag >> 2) + 1;
            off = b1 | (C.byte(ip + 2) << 8) | (C.byte(ip + 3) << 16) | (C.byte(ip + 4) << 24);
            adv = 5;
        }
        if (off == 0 || off > (uint32_t) P.op || P.op + (int32_t) ml + 8 > C.out_cap) return lzs::kFallback;
        if (!C.emit(ip, 0, ml, off)) return lzs::kFallback;
        P.ip = ip + adv;
        P.op += (int32_t) ml;
        return lzs::kMore;
    }

    // the step decoder resumes at the element at (ip, op); writes out_len / status
    static __device__ __forceinline__ void finish(const AccBatch &b, uint32_t blk, uint32_t ip, uint32_t op, int lane)
    {
        const uint8_t *in = b.src + b.src_off[blk];
        uint8_t *out = b.dst + b.dst_off[blk];
        const int64_t in_len = b.src_len[blk], out_cap = b.dst_cap[blk];
        if (ip == lzs::kWholeBlock) snappy_decode_block<true, true>(in, in_len, out, out_cap, b.out_len + blk, b.status + blk, lane);
        else snappy_decode_block<true, true>(in, in_len, out, out_cap, b.out_len + blk, b.status + blk, lane, (int64_t) ip, (int64_t) op);
    }
};

}  // namespace snappydec
