// lz4.cu -- LZ4 block decode / encode kernels for sm_100a.
//
// Replaces the reference's Lz4RawDecompressor.decompress (lz4/Lz4RawDecompressor.java:35-198) and
// Lz4RawCompressor.compress (lz4/Lz4RawCompressor.java:69-192).  Decode is bit-exact with the Java
// decoder, including which streams it rejects and the offset it reports; encode emits a valid LZ4
// block the Java decoder accepts (round-trip parity, the reference's own contract:
// AbstractTestCompression.java:362-393).
#include "acc_device.cuh"
#include "lz4_decode_v1.cuh"
#include "lz4_records.cuh"

namespace {

using namespace lz4v1;

// kMinCtas: resident CTAs per SM the register allocation is bounded for (1 = compiler's choice)
template <int kFast, int kMinCtas>
__global__ void __launch_bounds__(256, kMinCtas) lz4_decompress_kernel(AccBatch b)
{
    const int lane = lane_id();
    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;
        lz4_decode_block<kFast>(b.src + b.src_off[idx], b.src_len[idx], b.dst + b.dst_off[idx], b.dst_cap[idx],
                                b.out_len, b.status, idx, lane);
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

// lanes cooperatively write a sequence's token (literal length nibble + match code nibble, known together: the token is
// written once, never read back), the literal-length extension bytes and the literals; returns the new position.
// mcode < 0: last literals (no match part).
__device__ __forceinline__ int64_t lz4_emit_literal_run(uint8_t *out, int64_t op, const uint8_t *lit, int64_t ll, int64_t mcode, int lane)
{
    const uint32_t ml_nibble = mcode < 0 ? 0u : (mcode >= 15 ? 15u : (uint32_t) mcode);
    uint8_t *token = out + op++;
    if (ll >= 15) {
        if (lane == 0) *token = (uint8_t) (0xF0u | ml_nibble);
        int64_t rem = ll - 15;
        int64_t n255 = rem / 255;
        for (int64_t i = lane; i < n255; i += 32) out[op + i] = 255;
        if (lane == 0) out[op + n255] = (uint8_t) (rem - n255 * 255);
        op += n255 + 1;
    }
    else {
        if (lane == 0) *token = (uint8_t) ((uint32_t) (ll << 4) | ml_nibble);
    }
    warp_copy(out + op, lit, ll, lane);
    return op + ll;
}

// TableT = uint16_t serves blocks of at most 64 KiB (positions fit 16 bits, 8 KiB of table per warp -> twice the resident
// warps); TableT = int32_t serves larger blocks.  Each instantiation skips the blocks that belong to the other one.
template <typename TableT>
__global__ void __launch_bounds__(kLz4WarpsPerCta * 32) lz4_compress_kernel(AccBatch b)
{
    extern __shared__ __align__(16) uint8_t lz4_tables_raw[];  // kLz4WarpsPerCta x kLz4Table positions
    TableT *lz4_tables = reinterpret_cast<TableT *>(lz4_tables_raw);
    constexpr bool kSmallBlocks = sizeof(TableT) == 2;
    constexpr TableT kEmpty = (TableT) -1;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    TableT *table = lz4_tables + warp * kLz4Table;

    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;

        const uint8_t *in = b.src + b.src_off[idx];
        const int64_t in_len = b.src_len[idx];
        uint8_t *out = b.dst + b.dst_off[idx];
        const int64_t out_cap = b.dst_cap[idx];

        if ((in_len <= 65536) != kSmallBlocks) continue;   // handled by the other instantiation
        if (in_len > 0x7E000000) {
            if (lane == 0) { b.out_len[idx] = 0; b.status[idx] = ACC_STATUS(ACC_E_ARGUMENT, ACC_R_MAX_INPUT_EXCEEDED); }
            continue;
        }
        if (out_cap < in_len + in_len / 255 + 16) {
            if (lane == 0) { b.out_len[idx] = 0; b.status[idx] = ACC_STATUS(ACC_E_ARGUMENT, ACC_R_MAX_OUTPUT_TOO_SMALL); }
            continue;
        }

        for (int i = lane; i < kLz4Table; i += 32) table[i] = kEmpty;
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
                const bool live = p <= match_find_limit;
                bool hit = false;
                int32_t cand = -1;
                uint32_t slot = 0xFFFFFFFFu - (uint32_t) lane;          // idle lanes: distinct dummies for the insert below
                if (live) {
                    const uint64_t v = ld_u64_unaligned(in + p);
                    slot = lz4_hash5(v);
                    const TableT tv = table[slot];
                    cand = tv == kEmpty ? -1 : (int32_t) tv;
                    if (cand >= 0 && cand < p && p - cand <= 65535 && ld_u32_unaligned(in + cand) == (uint32_t) v) hit = true;
                }
                __syncwarp();
                // Insert after the lookups (lanes see the table as of the batch start), and only positions up to
                // the first match: a position behind the match end would otherwise be looked up again by the next
                // batch and find itself instead of its older candidate.  (Two lanes of one step can hash to
                // the same slot; which store lands last is the hardware's choice -- in practice the highest lane, and
                // tests/test_gpu_determinism.py pins that the output does not vary -- a __match_any_sync arbitration made
                // the choice explicit but cost 15 % of the kernel.)
                unsigned hits = __ballot_sync(kFull, hit);
                const int first_hit = hits ? __ffs(hits) - 1 : 31;
                if (live && lane <= first_hit) table[slot] = (TableT) p;
                __syncwarp();                                            // the next step's lookups see these inserts
                if (hits == 0) {
                    pos += 32;
                    continue;
                }
                const int first = __ffs(hits) - 1;
                int64_t mpos = pos + first;                              // match start in input
                int64_t ref = __shfl_sync(kFull, cand, first);           // candidate position
                // catch up backwards (Lz4RawCompressor.java:141-144), 32 bytes per ballot instead of a chain of dependent byte loads
                for (;;) {
                    const int64_t room = (mpos - anchor) < ref ? (mpos - anchor) : ref;
                    const bool eq1 = lane < room && in[mpos - 1 - lane] == in[ref - 1 - lane];
                    const unsigned m = __ballot_sync(kFull, eq1);
                    const int nb = m == kFull ? 32 : __ffs(~m) - 1;
                    mpos -= nb; ref -= nb;
                    if (nb < 32) break;
                }

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

                // ---- emit: token + literals [anchor, mpos), then offset and match-length extension ----
                const uint32_t offset = (uint32_t) (mpos - ref);
                const int64_t mcode = mlen - kMinMatch;
                op = lz4_emit_literal_run(out, op, in + anchor, mpos - anchor, mcode, lane);
                if (lane == 0) {
                    out[op] = (uint8_t) offset;
                    out[op + 1] = (uint8_t) (offset >> 8);
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
        op = lz4_emit_literal_run(out, op, in + anchor, in_len - anchor, -1, lane);
        if (lane == 0) { b.out_len[idx] = op; b.status[idx] = 0; }
        __syncwarp();
    }
}


// ---- record path (lz_records.cuh): parse kernel (one lane per block) + execute kernel (one warp per block) ----
constexpr int kParseThreads = 256;

__global__ void __launch_bounds__(kParseThreads) lz4_parse_kernel(AccBatch b, uint2 *recs, lzs::RecHeader *hdrs, int row)
{
    __shared__ __align__(16) uint8_t win[kParseThreads * lzs::kWinStride];
    lzs::parse_lane<Lz4Records>(b, win + threadIdx.x * lzs::kWinStride, recs, hdrs, row);
}

__global__ void __launch_bounds__(256, 6) lz4_execute_kernel(AccBatch b, const uint2 *recs, const lzs::RecHeader *hdrs, int row)
{
    __shared__ __align__(16) uint8_t rings[8 * lzs::kOutRing];
    lzs::execute_warp<Lz4Records>(b, recs, hdrs, row, rings + (threadIdx.x >> 5) * lzs::kOutRing, lane_id());
}

}  // namespace

// scratch of the record path for a batch of n blocks: one header + one row of records per block.  Rows hold up to 32,768
// records and share about 4 GiB (a 64 KiB Silesia block has ~4,000 sequences; a block with more than its row holds
// resumes the step decoder there)
int64_t acc_lz_records_row(int64_t n)
{
    int64_t row = (4LL << 30) / 8 / (n > 0 ? n : 1);
    if (row > 32768) row = 32768;
    if (row < 256) row = 256;
    return row;
}
int64_t acc_lz_records_scratch_bytes(int64_t n) { return n * ((int64_t) sizeof(lzs::RecHeader) + acc_lz_records_row(n) * 8) + 256; }

namespace {
template <class ParseK, class ExecK>
void launch_record_path(ParseK parse_k, ExecK exec_k, const AccBatch &b, int sm_count, cudaStream_t st, void *scratch, unsigned int *second_counter)
{
    const int row = (int) acc_lz_records_row(b.n);
    lzs::RecHeader *hdrs = reinterpret_cast<lzs::RecHeader *>(scratch);
    uint2 *recs = reinterpret_cast<uint2 *>(hdrs + b.n);
    // parse: one thread per block, all blocks at once when they fit the machine (2048 threads per SM)
    int64_t pctas = (b.n + kParseThreads - 1) / kParseThreads;
    const int64_t pmax = (int64_t) sm_count * (2048 / kParseThreads);
    if (pctas > pmax) pctas = pmax;
    if (pctas < 1) pctas = 1;
    parse_k<<<(unsigned) pctas, kParseThreads, 0, st>>>(b, recs, hdrs, row);
    // execute: one warp per block, persistent
    AccBatch b2 = b;
    b2.work_counter = second_counter;
    int64_t ectas = (b.n + 7) / 8;
    const int64_t emax = (int64_t) sm_count * 6;
    if (ectas > emax) ectas = emax;
    if (ectas < 1) ectas = 1;
    exec_k<<<(unsigned) ectas, 256, 0, st>>>(b2, recs, hdrs, row);
}
}  // namespace

void acc_launch_lz4_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st, void *scratch, unsigned int *second_counter)
{
    if (scratch) { launch_record_path(lz4_parse_kernel, lz4_execute_kernel, b, sm_count, st, scratch, second_counter); return; }
    // the step decoder alone (acc_set_tuning key 1): multi-sequence + medium steps, registers bounded for 8 resident CTAs
    // (32 registers, 64 warps per SM)
    if (ctas_per_sm <= 0) ctas_per_sm = 8;
    int64_t ctas = (b.n + 7) / 8;
    int64_t max_ctas = (int64_t) sm_count * ctas_per_sm;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    lz4_decompress_kernel<3, 8><<<(unsigned) ctas, 256, 0, st>>>(b);
}

void acc_launch_lz4_compress(const AccBatch &b, int sm_count, cudaStream_t st, unsigned int *second_counter)
{
    int64_t ctas = (b.n + kLz4WarpsPerCta - 1) / kLz4WarpsPerCta;
    // blocks <= 64 KiB: 16-bit tables, 32 KiB per CTA -> 7 CTAs (28 warps) per SM
    {
        const int smem = kLz4WarpsPerCta * kLz4Table * (int) sizeof(uint16_t);
        cudaFuncSetAttribute(lz4_compress_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int64_t g = ctas < (int64_t) sm_count * 7 ? ctas : (int64_t) sm_count * 7;
        if (g < 1) g = 1;
        lz4_compress_kernel<uint16_t><<<(unsigned) g, kLz4WarpsPerCta * 32, smem, st>>>(b);
    }
    // larger blocks: 32-bit tables (64 KiB per CTA -> 3 CTAs per SM); exits immediately when there are none
    {
        AccBatch b2 = b;
        b2.work_counter = second_counter;
        const int smem = kLz4WarpsPerCta * kLz4Table * (int) sizeof(int32_t);
        cudaFuncSetAttribute(lz4_compress_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int64_t g = ctas < (int64_t) sm_count * 3 ? ctas : (int64_t) sm_count * 3;
        if (g < 1) g = 1;
        lz4_compress_kernel<int32_t><<<(unsigned) g, kLz4WarpsPerCta * 32, smem, st>>>(b2);
    }
}
