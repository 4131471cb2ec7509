// lz4_records.cuh -- the LZ4 parse side of the record path (lz_records.cuh).  Included by lz4.cu and by the host emulation
// (tests/host/lzs_emu.cpp).
#pragma once
#include "lz4_decode_v1.cuh"
#include "lz_records.cuh"

namespace lz4v1 {

// parse_run walks this lane's block until it leaves the fast path (kFallback) or its record row is full.  A sequence is taken only under the conditions of the
// Java decoder's normal path (Lz4RawDecompressor.java:66-96: literals end >= 8 bytes before the input end and >= 12
// bytes before the output end; :116-119: 1 <= offset <= output position; :126-128,168: match ends >= 12 bytes before the
// output end), with every byte it reads at least 16 bytes before the input end so that the bounds of the Java length
// loops (:69-73 `input < inputLimit - 15`, :126 `input > inputLimit - 5`) cannot trigger.  Anything else returns
// kFallback at the token, where the step decoder of lz4_decode_v1.cuh takes over.
struct Lz4Records {
    struct Parse {
        int32_t ip, op;            // next token, next output byte
        int32_t tok_ip, tok_op;    // the sequence being parsed (restart point of a fallback)
        int32_t mpos;              // position of its offset bytes
        int32_t lit_pos, lit_rem;  // mode 1: rest of a long literal run, handed over in pieces
        uint32_t pll;              // mode 2: literals that ride in the match record (runs below 15 bytes)
        uint32_t tok_ml;
        uint32_t mode;             // 0 token, 1 long literal run, 2 offset + match length
    };
    static __device__ __forceinline__ void begin(Parse &P) { P.ip = 0; P.op = 0; P.tok_ip = 0; P.tok_op = 0; P.mode = 0; P.pll = 0; P.lit_rem = 0; P.lit_pos = 0; P.mpos = 0; P.tok_ml = 0; }
    // where the step decoder takes over: the token of the sequence that was not (completely) recorded
    static __device__ __forceinline__ uint32_t resume_ip(const Parse &P) { return (uint32_t) P.tok_ip; }
    static __device__ __forceinline__ uint32_t resume_op(const Parse &P) { return (uint32_t) P.tok_op; }

    static __device__ __forceinline__ int parse_run(Parse &P, lzs::ParseCtx &C, const int budget)
    {
        const int32_t safe_end = C.in_len - 16;      // bytes at positions < safe_end may be read without a bounds story
        while (C.n_rec < budget) {
            if (P.mode == 0) {
                const int32_t ip = P.ip;
                P.tok_ip = ip; P.tok_op = P.op;
                if (ip >= safe_end) return lzs::kFallback;
                C.ensure(ip);                                // token, a short literal run's offset and one extension byte lie within 32 bytes
                const uint32_t tok = C.byte(ip);
                uint32_t ll = tok >> 4;
                int32_t p = ip + 1;
                if (ll == 15) {
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
                P.tok_ml = tok & 15;
                P.mpos = p + (int32_t) ll;
                if (ll < 15) { P.pll = ll; P.mode = 2; }
                else { P.pll = 0; P.lit_pos = p; P.lit_rem = (int32_t) ll; P.mode = 1; }
            }
            if (P.mode == 1) {
                const int32_t n = P.lit_rem < lzs::kMaxLitPiece ? P.lit_rem : lzs::kMaxLitPiece;
                if (!C.emit(P.lit_pos, (uint32_t) n, 0, lzs::kNoOffset)) return lzs::kFallback;
                P.lit_pos += n; P.lit_rem -= n; P.op += n;
                if (P.lit_rem == 0) P.mode = 2;
                continue;
            }
            // mode 2: offset, match length
            const int32_t mpos = P.mpos;
            if (mpos + 2 > safe_end) return lzs::kFallback;
            if (P.pll == 0) C.ensure(mpos);                  // behind a long literal run; otherwise still inside the token's 32 bytes
            const uint32_t off = C.byte(mpos) | (C.byte(mpos + 1) << 8);
            uint32_t ml = P.tok_ml;
            int32_t p2 = mpos + 2;
            if (ml == 15) {
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
            const int32_t mop = P.op + (int32_t) P.pll;                 // output position of the match
            if (off == 0 || (int32_t) off > mop || mop + (int32_t) ml + 12 > C.out_cap) return lzs::kFallback;
            if (!C.emit(mpos - (int32_t) P.pll, P.pll, ml, off)) return lzs::kFallback;
            P.ip = p2;
            P.op = mop + (int32_t) ml;
            P.mode = 0;
        }
        // row full: between two sequences the next token is the resume point; inside one, its own token (the literal pieces
        // already recorded are written again by the step decoder: same bytes)
        if (P.mode == 0) { P.tok_ip = P.ip; P.tok_op = P.op; }
        return lzs::kRowFull;
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
