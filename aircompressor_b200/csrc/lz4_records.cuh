// lz4_records.cuh -- the LZ4 parse side of the record path (lz_records.cuh).  Included by lz4.cu and by the host emulation
// (tests/host/lzs_emu.cpp).
#pragma once
#include "lz4_decode_v1.cuh"
#include "lz_records.cuh"

namespace lz4v1 {

// parse_one records one sequence of this lane's block, or says that the block leaves the fast path here.  A sequence is taken only under the conditions of the
// Java decoder's normal path (Lz4RawDecompressor.java:66-96: literals end >= 8 bytes before the input end and >= 12
// bytes before the output end; :116-119: 1 <= offset <= output position; :126-128,168: match ends >= 12 bytes before the
// output end), with every byte it reads at least 16 bytes before the input end so that the bounds of the Java length
// loops (:69-73 `input < inputLimit - 15`, :126 `input > inputLimit - 5`) cannot trigger.  Anything else returns
// kFallback at the token, where the step decoder of lz4_decode_v1.cuh takes over.
struct Lz4Records {
    struct Parse {
        int32_t ip, op;            // next token, next output byte
        int32_t tok_ip, tok_op;    // where the step decoder takes over: the token of the sequence that was not recorded
    };
    static __device__ __forceinline__ void begin(Parse &P) { P.ip = 0; P.op = 0; P.tok_ip = 0; P.tok_op = 0; }
    static __device__ __forceinline__ uint32_t resume_ip(const Parse &P) { return (uint32_t) P.tok_ip; }
    static __device__ __forceinline__ uint32_t resume_op(const Parse &P) { return (uint32_t) P.tok_op; }

    // One sequence: token, literal length (+ extension), offset, match length (+ extension), all checks, then its record(s).
    // The common case -- at most one extension byte per length -- is straight-line code; nothing is recorded before the
    // sequence is known to be valid, so a hand-over always happens at a token.
    static __device__ __forceinline__ int parse_one(Parse &P, lzs::ParseCtx &C, const int row)
    {
        const int32_t safe_end = C.in_len - 16;      // bytes at positions < safe_end may be read without a bounds story
        const int32_t ip = P.ip;
        P.tok_ip = ip; P.tok_op = P.op;
        if (C.n_rec + 2 > row) return lzs::kRowFull;
        if (ip + 20 > safe_end) return lzs::kFallback;   // token + extension + 14 literals + offset + extension of a short sequence: all readable
        C.ensure(ip);                                // ... and all within the 32 bytes the window guarantees
        const uint32_t tok = C.byte(ip);
        // the first extension byte of either length is read whether the token asks for it or not: the common sequence (at most
        // one extension byte per length) is straight-line code, which is what keeps the 32 lanes of the warp together
        const bool lx = tok >= 0xF0;
        const uint32_t e1 = C.byte(ip + 1);
        uint32_t ll = (tok >> 4) + (lx ? e1 : 0u);
        int32_t p = ip + 1 + (lx ? 1 : 0);
        if (lx && e1 == 255) {                       // rare: further extension bytes
            uint32_t v;
            int cnt = 0;
            do {
                if (p >= safe_end || ++cnt > 64) return lzs::kFallback;
                C.ensure(p);
                v = C.byte(p++);
                ll += v;
            }
            while (v == 255);
        }
        if (p + (int32_t) ll + 8 > C.in_len || P.op + (int32_t) ll + 12 > C.out_cap) return lzs::kFallback;
        const int32_t mpos = p + (int32_t) ll;
        if (mpos + 4 > safe_end) return lzs::kFallback;
        if (lx) C.ensure(mpos);                      // behind a long literal run the window moves on
        const uint32_t off = C.byte(mpos) | (C.byte(mpos + 1) << 8);
        const bool mx = (tok & 15) == 15;
        const uint32_t e2 = C.byte(mpos + 2);
        uint32_t ml = (tok & 15) + (mx ? e2 : 0u);
        int32_t p2 = mpos + 2 + (mx ? 1 : 0);
        if (mx && e2 == 255) {                       // rare
            uint32_t v;
            do {
                if (p2 >= safe_end || ml > (1u << 19)) return lzs::kFallback;
                C.ensure(p2);
                v = C.byte(p2++);
                ml += v;
            }
            while (v == 255);
        }
        ml += kMinMatch;
        const int32_t mop = P.op + (int32_t) ll;     // output position of the match
        if (off == 0 || (int32_t) off > mop || mop + (int32_t) ml + 12 > C.out_cap) return lzs::kFallback;
        if (lx && ll > (uint32_t) lzs::kMaxLitPiece) {
            // rare: a literal run beyond the 12-bit length field travels as several literal-only records
            int32_t lp = p, rem = (int32_t) ll;
            while (rem > lzs::kMaxLitPiece) {
                if (C.n_rec + 3 > row) return lzs::kRowFull;           // the pieces already recorded are written again by the step decoder: same bytes
                if (!C.emit(lp, (uint32_t) lzs::kMaxLitPiece, 0, lzs::kNoOffset)) return lzs::kFallback;
                lp += lzs::kMaxLitPiece; rem -= lzs::kMaxLitPiece;
            }
            p = lp; ll = (uint32_t) rem;
        }
        // one record, or (long literal run) a literal-only record and the match on its own
        if (!C.emit2(p, ll, ml, off, lx, mpos)) return lzs::kFallback;
        P.ip = p2;
        P.op = mop + (int32_t) ml;
        return lzs::kMore;
    }

    // the step decoder of lz4_decode_v1.cuh finishes the block from the token at (ip, op); writes out_len / status
    static __device__ __forceinline__ void finish(const AccBatch &b, uint32_t blk, uint32_t ip, uint32_t op, int lane)
    {
        const uint8_t *in = b.src + b.src_off[blk];
        uint8_t *out = b.dst + b.dst_off[blk];
        const int64_t in_len = b.src_len[blk], out_cap = b.dst_cap[blk];
        if (ip == lzs::kWholeBlock) lz4_decode_block<3>(in, in_len, out, out_cap, b.out_len, b.status, blk, lane);
        else lz4_decode_impl<3, int32_t>(in, (int32_t) in_len, out, (int32_t) out_cap, b.out_len, b.status, blk, lane, (int32_t) ip, (int32_t) op);
    }
};

}  // namespace lz4v1
