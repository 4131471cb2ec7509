// snappy.cu -- Snappy raw-format decode / encode kernels for sm_100a.
//
// Replaces SnappyRawDecompressor.decompress / uncompressAll (snappy/SnappyRawDecompressor.java:35-220)
// and SnappyRawCompressor.compress (snappy/SnappyRawCompressor.java:74-233).  Decode is bit-exact with
// the Java decoder (same accept/reject decisions and error offsets); encode emits a valid stream the
// Java decoder round-trips (AbstractTestCompression.java:362-393).
#include "acc_device.cuh"
#include "snappy_decode.cuh"

namespace {

using namespace snappydec;

template <bool kMulti, int kMinCtas>
__global__ void __launch_bounds__(256, kMinCtas) snappy_decompress_kernel(AccBatch b)
{
    const int lane = lane_id();
    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;
        snappy_decode_block<kMulti>(b.src + b.src_off[idx], b.src_len[idx], b.dst + b.dst_off[idx], b.dst_cap[idx],
                            b.out_len + idx, b.status + idx, lane);
    }
}


// ------------------------------------------------------------------------------------------------
// Encode: one warp per input; independent 64 KiB fragments (SnappyRawCompressor.java:37-38,93) are
// walked in order with an 8,192 x uint16 position table in shared memory (half the size of the
// reference's short[16384], SnappyRawCompressor.java:43-45,348-361) that is reset per fragment.
// 32 positions are probed per step with the reference's hash (value * 0x1e35a7bd >>> shift,
// SnappyRawCompressor.java:368-371); ballot selects the first match, a second ballot extends it.
// ------------------------------------------------------------------------------------------------
constexpr int kSnTableBits = 13;   // 8192 x uint16 per warp (the reference uses 16384): 16 KiB per warp, twice the resident warps
constexpr int kSnTable = 1 << kSnTableBits;
constexpr int kSnWarpsPerCta = 4;
constexpr int kSnFragment = 1 << 16;

__device__ __forceinline__ uint32_t snappy_hash(uint32_t v) { return (v * 0x1e35a7bdu) >> (32 - kSnTableBits); }

__device__ __forceinline__ int64_t snappy_emit_literal(uint8_t *out, int64_t op, const uint8_t *lit, int64_t ll, int lane)
{
    // SnappyRawCompressor.java:268-298 emitLiteralLength
    uint32_t n = (uint32_t) (ll - 1);
    int hdr;
    if (n < 60) { hdr = 1; if (lane == 0) out[op] = (uint8_t) (n << 2); }
    else {
        int bytes = n < (1u << 8) ? 1 : n < (1u << 16) ? 2 : n < (1u << 24) ? 3 : 4;
        hdr = 1 + bytes;
        if (lane == 0) out[op] = (uint8_t) ((59 + bytes) << 2);
        if (lane >= 1 && lane <= bytes) out[op + lane] = (uint8_t) (n >> (8 * (lane - 1)));
    }
    op += hdr;
    warp_copy(out + op, lit, ll, lane);
    return op + ll;
}

__device__ __forceinline__ int64_t snappy_emit_copy(uint8_t *out, int64_t op, uint32_t offset, int64_t len, int lane)
{
    // SnappyRawCompressor.java:312-345 emitCopy: 64-byte COPY_2 chunks while len >= 68, one 60-byte
    // chunk if len > 64, then COPY_1 (len 4..11, offset < 2048) or COPY_2.
    int64_t n64 = len >= 68 ? (len - 68) / 64 + 1 : 0;
    for (int64_t i = lane; i < n64; i += 32) {
        uint8_t *o = out + op + i * 3;
        o[0] = (uint8_t) (2 + ((64 - 1) << 2)); o[1] = (uint8_t) offset; o[2] = (uint8_t) (offset >> 8);
    }
    op += n64 * 3;
    len -= n64 * 64;
    if (lane == 0) {
        if (len > 64) {
            out[op] = (uint8_t) (2 + ((60 - 1) << 2)); out[op + 1] = (uint8_t) offset; out[op + 2] = (uint8_t) (offset >> 8);
        }
    }
    if (len > 64) { op += 3; len -= 60; }
    if (len < 12 && offset < 2048) {
        if (lane == 0) {
            out[op] = (uint8_t) (1 + ((len - 4) << 2) + ((offset >> 8) << 5));
            out[op + 1] = (uint8_t) offset;
        }
        op += 2;
    }
    else {
        if (lane == 0) {
            out[op] = (uint8_t) (2 + ((len - 1) << 2)); out[op + 1] = (uint8_t) offset; out[op + 2] = (uint8_t) (offset >> 8);
        }
        op += 3;
    }
    return op;
}

__global__ void __launch_bounds__(kSnWarpsPerCta * 32) snappy_compress_kernel(AccBatch b)
{
    extern __shared__ uint16_t sn_tables[];  // kSnWarpsPerCta x kSnTable
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    uint16_t *table = sn_tables + warp * kSnTable;

    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;

        const uint8_t *in = b.src + b.src_off[idx];
        const int64_t in_len = b.src_len[idx];
        uint8_t *out = b.dst + b.dst_off[idx];
        const int64_t out_cap = b.dst_cap[idx];
        if (in_len > 0x7fffffff || out_cap < 32 + in_len + in_len / 6) {
            if (lane == 0) { b.out_len[idx] = 0; b.status[idx] = ACC_STATUS(ACC_E_ARGUMENT, ACC_R_MAX_OUTPUT_TOO_SMALL); }
            continue;
        }
        // preamble: varint(uncompressed length) (SnappyRawCompressor.java:383-411)
        int64_t op = 0;
        {
            uint32_t v = (uint32_t) in_len;
            int nb = v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5;
            if (lane < nb) out[lane] = (uint8_t) (((v >> (7 * lane)) & 0x7f) | (lane < nb - 1 ? 0x80 : 0));
            op = nb;
        }

        for (int64_t frag = 0; frag < in_len; frag += kSnFragment) {
            const int64_t frag_limit = (in_len < frag + kSnFragment) ? in_len : frag + kSnFragment;
            uint32_t *t32 = (uint32_t *) table;
            for (int i = lane; i < kSnTable / 2; i += 32) t32[i] = 0xffffffffu;  // 0xffff = empty
            __syncwarp();
            const uint8_t *base = in + frag;
            const int64_t flen = frag_limit - frag;
            // keep the reference's margin: no match starts within the last 15 bytes of a fragment
            const int64_t fast_limit = flen - 15;
            int64_t next_emit = 0;
            int64_t pos = 0;
            while (pos <= fast_limit) {
                const int64_t p = pos + lane;
                const bool live = p <= fast_limit;
                bool hit = false;
                uint32_t cand = 0xffff;
                uint32_t slot = 0xFFFFFFFFu - (uint32_t) lane;          // idle lanes: distinct dummies for the insert below
                if (live) {
                    const uint32_t cur = ld_u32_unaligned(base + p);
                    slot = snappy_hash(cur);
                    cand = table[slot];
                    if (cand != 0xffff && cand < p && ld_u32_unaligned(base + cand) == cur) hit = true;
                }
                __syncwarp();
                // insert only up to the first match (see lz4.cu); position 65535
                // cannot be stored (0xffff = empty)
                unsigned hits = __ballot_sync(kFull, hit);
                const int first_hit = hits ? __ffs(hits) - 1 : 31;
                if (live && p < 0xffff && lane <= first_hit) table[slot] = (uint16_t) p;
                __syncwarp();                                            // the next step's lookups see these inserts
                if (hits == 0) { pos += 32; continue; }
                const int first = __ffs(hits) - 1;
                const int64_t mpos = pos + first;
                const int64_t ref = __shfl_sync(kFull, cand, first);
                int64_t mlen = 4;
                for (;;) {
                    int64_t q = mpos + mlen + lane;
                    bool same = (q < flen) && (base[q] == base[ref + mlen + lane]);
                    unsigned eq = __ballot_sync(kFull, same);
                    if (eq == kFull) { mlen += 32; continue; }
                    mlen += __ffs(~eq) - 1;
                    break;
                }
                if (mpos > next_emit) op = snappy_emit_literal(out, op, base + next_emit, mpos - next_emit, lane);
                op = snappy_emit_copy(out, op, (uint32_t) (mpos - ref), mlen, lane);
                pos = mpos + mlen;
                next_emit = pos;
                __syncwarp();
            }
            if (next_emit < flen) op = snappy_emit_literal(out, op, base + next_emit, flen - next_emit, lane);
            __syncwarp();
        }
        if (lane == 0) { b.out_len[idx] = op; b.status[idx] = 0; }
        __syncwarp();
    }
}


// ---- record path (lz_records.cuh) ----
constexpr int kParseThreads = 256;

__global__ void __launch_bounds__(kParseThreads) snappy_parse_kernel(AccBatch b, uint2 *recs, lzs::RecHeader *hdrs, int row)
{
    __shared__ __align__(16) uint8_t win[kParseThreads * lzs::kWinStride];
    lzs::parse_lane<SnappyRecords>(b, win + threadIdx.x * lzs::kWinStride, recs, hdrs, row);
}

__global__ void __launch_bounds__(256, 5) snappy_execute_kernel(AccBatch b, const uint2 *recs, const lzs::RecHeader *hdrs, int row)
{
    __shared__ __align__(16) uint8_t rings[8 * lzs::kOutRing];
    lzs::execute_warp<SnappyRecords>(b, recs, hdrs, row, rings + (threadIdx.x >> 5) * lzs::kOutRing, lane_id());
}

}  // namespace

int64_t acc_lz_records_row(int64_t n);   // lz4.cu

void acc_launch_snappy_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st, void *scratch, unsigned int *second_counter)
{
    if (scratch) {
        const int row = (int) acc_lz_records_row(b.n);
        lzs::RecHeader *hdrs = reinterpret_cast<lzs::RecHeader *>(scratch);
        uint2 *recs = reinterpret_cast<uint2 *>(hdrs + b.n);
        int64_t pctas = (b.n + kParseThreads - 1) / kParseThreads;
        const int64_t pmax = (int64_t) sm_count * (2048 / kParseThreads);
        if (pctas > pmax) pctas = pmax;
        if (pctas < 1) pctas = 1;
        snappy_parse_kernel<<<(unsigned) pctas, kParseThreads, 0, st>>>(b, recs, hdrs, row);
        AccBatch b2 = b;
        b2.work_counter = second_counter;
        int64_t ectas = (b.n + 7) / 8;
        const int64_t emax = (int64_t) sm_count * 5;
        if (ectas > emax) ectas = emax;
        if (ectas < 1) ectas = 1;
        snappy_execute_kernel<<<(unsigned) ectas, 256, 0, st>>>(b2, recs, hdrs, row);
        return;
    }
    // the step decoder alone (acc_set_tuning key 1)
    if (ctas_per_sm <= 0) ctas_per_sm = 8;
    int64_t ctas = (b.n + 7) / 8;
    int64_t max_ctas = (int64_t) sm_count * ctas_per_sm;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    snappy_decompress_kernel<true, 5><<<(unsigned) ctas, 256, 0, st>>>(b);
}

void acc_launch_snappy_compress(const AccBatch &b, int sm_count, cudaStream_t st)
{
    const int smem = kSnWarpsPerCta * kSnTable * (int) sizeof(uint16_t);
    cudaFuncSetAttribute(snappy_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  // per device, cheap
    int64_t ctas = (b.n + kSnWarpsPerCta - 1) / kSnWarpsPerCta;
    int64_t max_ctas = (int64_t) sm_count * 3;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    snappy_compress_kernel<<<(unsigned) ctas, kSnWarpsPerCta * 32, smem, st>>>(b);
}
