// lz4.cu -- LZ4 block decode / encode kernels for sm_100a.
//
// Replaces the reference's Lz4RawDecompressor.decompress (lz4/Lz4RawDecompressor.java:35-198) and
// Lz4RawCompressor.compress (lz4/Lz4RawCompressor.java:69-192).  Decode is bit-exact with the Java
// decoder, including which streams it rejects and the offset it reports; encode emits a valid LZ4
// block the Java decoder accepts (round-trip parity, the reference's own contract:
// AbstractTestCompression.java:362-393).
#include "acc_device.cuh"

namespace {

constexpr int kMinMatch = 4;
constexpr int kLastLiterals = 5;

// floor(65536 / d) + 1: (m * kRcp16[d]) >> 16 == m / d for m < 32
__constant__ uint32_t kRcp16[32] = {0, 65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554, 5958, 5462, 5042, 4682, 4370,
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
        if (small && ip + 25 <= in_len && op + 44 <= out_cap) {
            const uint32_t tk = __ldg(in + ip);
            const uint32_t fll = tk >> 4, fml = tk & 15;
            if (fll != 15 && fml != 15) {
                const uint32_t lit0 = (uint32_t) ip + 1;
                const uint32_t foff = (uint32_t) __ldg(in + lit0 + fll) | ((uint32_t) __ldg(in + lit0 + fll + 1) << 8);
                if (foff == 0 || foff > (uint32_t) op + fll) LZ4_FAIL((int64_t) lit0 + fll + 2, ACC_R_OFFSET_OUTSIDE);
                const uint32_t total = fll + fml + kMinMatch;
                if ((uint32_t) lane < total) {
                    uint8_t v;
                    if ((uint32_t) lane < fll) v = __ldg(in + lit0 + lane);
                    else {
                        uint32_t m = (uint32_t) lane - fll;
                        if (m >= foff) m -= foff * ((m * kRcp16[foff]) >> 16);   // m % offset (offset < 18 here)
                        const int32_t rel = (int32_t) fll - (int32_t) foff + (int32_t) m;   // relative to op
                        v = rel >= 0 ? __ldg(in + lit0 + rel) : out[op + rel];
                    }
                    out[op + lane] = v;
                }
                __syncwarp();
                ip = lit0 + fll + 2;
                op += total;
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

__global__ void __launch_bounds__(256) lz4_decompress_kernel(AccBatch b)
{
    const int lane = lane_id();
    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;
        lz4_decode_block(b.src + b.src_off[idx], b.src_len[idx], b.dst + b.dst_off[idx], b.dst_cap[idx],
                         b.out_len + idx, b.status + idx, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Encode: one warp per block, 4096-entry position table in shared memory (same size as the
// reference's int[4096], Lz4RawCompressor.java:29-32,304-311).  The warp probes 32 consecutive
// positions at once: every lane hashes its own position with the reference's 5-byte multiplicative
// hash (Lz4RawCompressor.java:50-62), reads the candidate, verifies 4 bytes and the 64 KiB distance
// limit; a ballot picks the first lane with a match.  Match length is extended 32 bytes per step with
// a second ballot.  All end-of-block rules of the format (no match starts in the last 12 bytes, last
// 5 bytes are literals) are kept so the Java decoder accepts the stream with an exact-size output.
// ------------------------------------------------------------------------------------------------
constexpr int kLz4HashLog = 12;
constexpr int kLz4Table = 1 << kLz4HashLog;
constexpr int kLz4WarpsPerCta = 4;

__device__ __forceinline__ uint32_t lz4_hash5(uint64_t v)
{
    return (uint32_t) ((v * 889523592379ULL) >> 28) & (kLz4Table - 1);
}

// lanes cooperatively write the run-length header [token + extension bytes]; returns new position
__device__ __forceinline__ int64_t lz4_emit_literal_run(uint8_t *out, int64_t op, const uint8_t *lit, int64_t ll, uint8_t **token_ptr, int lane)
{
    uint8_t *token = out + op++;
    if (ll >= 15) {
        if (lane == 0) *token = 0xF0;
        int64_t rem = ll - 15;
        int64_t n255 = rem / 255;
        for (int64_t i = lane; i < n255; i += 32) out[op + i] = 255;
        if (lane == 0) out[op + n255] = (uint8_t) (rem - n255 * 255);
        op += n255 + 1;
    }
    else {
        if (lane == 0) *token = (uint8_t) (ll << 4);
    }
    warp_copy(out + op, lit, ll, lane);
    *token_ptr = token;
    return op + ll;
}

__global__ void __launch_bounds__(kLz4WarpsPerCta * 32) lz4_compress_kernel(AccBatch b)
{
    extern __shared__ int32_t lz4_tables[];  // kLz4WarpsPerCta x kLz4Table positions
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    int32_t *table = lz4_tables + warp * kLz4Table;

    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;

        const uint8_t *in = b.src + b.src_off[idx];
        const int64_t in_len = b.src_len[idx];
        uint8_t *out = b.dst + b.dst_off[idx];
        const int64_t out_cap = b.dst_cap[idx];

        if (in_len > 0x7E000000) {
            if (lane == 0) { b.out_len[idx] = 0; b.status[idx] = ACC_STATUS(ACC_E_ARGUMENT, ACC_R_MAX_INPUT_EXCEEDED); }
            continue;
        }
        if (out_cap < in_len + in_len / 255 + 16) {
            if (lane == 0) { b.out_len[idx] = 0; b.status[idx] = ACC_STATUS(ACC_E_ARGUMENT, ACC_R_MAX_OUTPUT_TOO_SMALL); }
            continue;
        }

        for (int i = lane; i < kLz4Table; i += 32) table[i] = -1;
        __syncwarp();

        int64_t op = 0;
        int64_t anchor = 0;
        const int64_t match_find_limit = in_len - 12;  // last position where a match may start
        const int64_t match_limit = in_len - kLastLiterals;
        int64_t pos = 0;

        if (in_len >= 13) {
            while (pos <= match_find_limit) {
                // ---- probe 32 positions ----
                const int64_t p = pos + lane;
                bool hit = false;
                int32_t cand = -1;
                if (p <= match_find_limit) {
                    uint64_t v = ld_u64_unaligned(in + p);
                    uint32_t h = lz4_hash5(v);
                    cand = table[h];
                    if (cand >= 0 && cand < p && p - cand <= 65535 && ld_u32_unaligned(in + cand) == (uint32_t) v) hit = true;
                }
                __syncwarp();
                // Insert after the lookups (lanes see the table as of the batch start), and only positions up to
                // the first match: a position behind the match end would otherwise be looked up again by the next
                // batch and find itself instead of its older candidate.
                unsigned hits = __ballot_sync(kFull, hit);
                const int first_hit = hits ? __ffs(hits) - 1 : 31;
                if (p <= match_find_limit && lane <= first_hit) {
                    uint64_t v = ld_u64_unaligned(in + p);
                    table[lz4_hash5(v)] = (int32_t) p;
                }
                if (hits == 0) {
                    pos += 32;
                    continue;
                }
                const int first = __ffs(hits) - 1;
                int64_t mpos = pos + first;                              // match start in input
                int64_t ref = __shfl_sync(kFull, cand, first);           // candidate position
                // catch up backwards (Lz4RawCompressor.java:141-144)
                while (mpos > anchor && ref > 0 && in[mpos - 1] == in[ref - 1]) { --mpos; --ref; }

                // ---- extend the match forwards, 32 bytes per ballot ----
                int64_t mlen = kMinMatch;
                for (;;) {
                    int64_t q = mpos + mlen + lane;
                    bool same = (q < match_limit) && (in[q] == in[ref + mlen + lane]);
                    unsigned eq = __ballot_sync(kFull, same);
                    if (eq == kFull) { mlen += 32; continue; }
                    mlen += __ffs(~eq) - 1;
                    break;
                }

                // ---- emit: literals [anchor, mpos), then the match ----
                uint8_t *token;
                op = lz4_emit_literal_run(out, op, in + anchor, mpos - anchor, &token, lane);
                const uint32_t offset = (uint32_t) (mpos - ref);
                int64_t mcode = mlen - kMinMatch;
                __syncwarp();
                if (lane == 0) {
                    out[op] = (uint8_t) offset;
                    out[op + 1] = (uint8_t) (offset >> 8);
                    if (mcode >= 15) *token |= 15; else *token |= (uint8_t) mcode;
                }
                op += 2;
                if (mcode >= 15) {
                    int64_t rem = mcode - 15;
                    int64_t n255 = rem / 255;
                    for (int64_t i = lane; i < n255; i += 32) out[op + i] = 255;
                    if (lane == 0) out[op + n255] = (uint8_t) (rem - n255 * 255);
                    op += n255 + 1;
                }
                pos = mpos + mlen;
                anchor = pos;
                __syncwarp();
            }
        }
        // ---- last literals ----
        uint8_t *token;
        op = lz4_emit_literal_run(out, op, in + anchor, in_len - anchor, &token, lane);
        if (lane == 0) { b.out_len[idx] = op; b.status[idx] = 0; }
        __syncwarp();
    }
}

}  // namespace

void acc_launch_lz4_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st)
{
    if (ctas_per_sm <= 0) ctas_per_sm = 8;
    int64_t warps_needed = b.n;
    int64_t ctas = (warps_needed + 7) / 8;
    int64_t max_ctas = (int64_t) sm_count * ctas_per_sm;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    lz4_decompress_kernel<<<(unsigned) ctas, 256, 0, st>>>(b);
}

void acc_launch_lz4_compress(const AccBatch &b, int sm_count, cudaStream_t st)
{
    int64_t ctas = (b.n + kLz4WarpsPerCta - 1) / kLz4WarpsPerCta;
    int64_t max_ctas = (int64_t) sm_count * 3;  // 64 KiB of tables per CTA -> 3 CTAs per SM
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    const int smem = kLz4WarpsPerCta * kLz4Table * (int) sizeof(int32_t);
    cudaFuncSetAttribute(lz4_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  // per device, cheap
    lz4_compress_kernel<<<(unsigned) ctas, kLz4WarpsPerCta * 32, smem, st>>>(b);
}
