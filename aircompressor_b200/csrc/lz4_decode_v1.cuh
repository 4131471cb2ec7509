// lz4_decode_v1.cuh -- exact warp-per-block LZ4 decoder (bit-exact with Lz4RawDecompressor.java:35-198, including
// every reject decision and error offset).  Used directly by lz4_decompress_kernel and as the fallback of the
// shared-memory-window decoder in lz4_v3.cu.
#pragma once
#include "acc_device.cuh"

namespace lz4v1 {

constexpr int kMinMatch = 4;
constexpr int kLastLiterals = 5;

// floor(65536 / d) + 1: (m * kRcp16[d]) >> 16 == m / d for m < 32
static __constant__ uint32_t kRcp16[32] = {0, 65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554, 5958, 5462, 5042, 4682, 4370,
                                    4097, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115};

// ------------------------------------------------------------------------------------------------
// Decode: one warp per block.  All lanes walk the token stream redundantly (broadcast loads), the
// literal and match copies are spread over the 32 lanes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lz4_decode_block(const uint8_t *__restrict__ in, int64_t in_len, uint8_t *out, int64_t out_cap,
                                                 int64_t *out_len, int32_t *status, int lane)
{
#define LZ4_FAIL(off, reason) do { if (lane == 0) { *out_len = (off); *status = ACC_STATUS(ACC_E_MALFORMED, reason); } return; } while (0)
    const int64_t fast_output_limit = out_cap - 8;
    int64_t ip = 0, op = 0;

    if (in_len == 0) LZ4_FAIL(0, ACC_R_INPUT_EMPTY);
    if (out_cap == 0) {
        if (in_len == 1 && in[0] == 0) { if (lane == 0) { *out_len = 0; *status = 0; } return; }
        if (lane == 0) { *out_len = 0; *status = ACC_STATUS(ACC_E_DST_TOO_SMALL, ACC_R_LZ4_ZERO_CAPACITY); }
        return;
    }

    // ---- fast path ------------------------------------------------------------------------------
    // Most sequences have no length-extension bytes (literal length < 15, match length < 19), i.e. they
    // produce at most 32 bytes.  For those, every lane resolves the source of ONE output byte directly
    // (a literal from the input, an older output byte, or -- when the match overlaps this sequence's own
    // literals / itself -- the literal it ultimately repeats) and the whole sequence is a single
    // load + store per lane.  The bounds below are exactly the conditions under which the Java decoder
    // takes its normal (non-final) path (Lz4RawDecompressor.java:82,168), so results are identical.
    const bool small = in_len < 0x7fffff00LL && out_cap < 0x7fffff00LL;
    while (ip < in_len) {
        if (small && ip + 32 <= in_len && op + 44 <= out_cap) {
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

        const int64_t lit_end = ip + (int64_t) ll;
        const int64_t lit_out_limit = op + (int64_t) ll;
        if (lit_out_limit > fast_output_limit - kMinMatch || lit_end > in_len - (2 + 1 + kLastLiterals)) {
            if (lit_out_limit > out_cap) LZ4_FAIL(ip, ACC_R_LAST_LITERAL_OUTSIDE);
            if (lit_end != in_len) LZ4_FAIL(ip, ACC_R_ALL_INPUT_CONSUMED);
            warp_copy(out + op, in + ip, ll, lane);
            op += ll;
            break;
        }
        warp_copy(out + op, in + ip, ll, lane);
        op = lit_out_limit;
        ip = lit_end;

        const uint32_t offset = ld_u16le(in + ip);
        ip += 2;
        if ((int64_t) offset > op || offset == 0) LZ4_FAIL(ip, ACC_R_OFFSET_OUTSIDE);

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

        const int64_t match_out_limit = op + (int64_t) ml;
        if (match_out_limit > fast_output_limit - kMinMatch) {
            if (match_out_limit > out_cap - kLastLiterals) LZ4_FAIL(ip, ACC_R_LAST5_LITERALS);
        }
        __syncwarp();
        warp_match_copy(out + op, offset, ml, lane);
        __syncwarp();
        op = match_out_limit;
    }
    if (lane == 0) { *out_len = op; *status = 0; }
#undef LZ4_FAIL
}


}  // namespace lz4v1
