// zstd_enc.cu -- Zstandard frame encode kernel for sm_100a.
//
// Replaces ZstdFrameCompressor.compress (zstd/ZstdFrameCompressor.java:136-260) and what it calls: the LZ77
// parse (zstd/DoubleFastBlockCompressor.java:28-180), the sequence store (zstd/SequenceStore.java), literal
// Huffman coding (zstd/HuffmanCompressionTable.java, zstd/HuffmanCompressor.java), sequence FSE coding
// (zstd/SequenceEncoder.java:228-297, zstd/FseCompressionTable.java) and the XXH64 frame checksum.
// Parity contract = the reference's own: the emitted frame must be accepted by the Java decoder rules
// (oracle) and by libzstd and reproduce the input (AbstractTestCompression.java:362-393); compressed bytes
// are not expected to be identical to the Java compressor's.
//
// Mapping: one 256-thread CTA per input.  Per block (<= 128 KiB):
//   1. match finding, the double-fast idea (DoubleFastBlockCompressor.java:28-180: a short and a long hash table, the
//      repeated offsets tried first) laid out for 8 warps that parse one eighth of the block each:
//        pass A: every warp fills a "long" table (6-byte hash, 2048 entries) with the LAST position of every value in
//                its own sub-range -- read-only afterwards, so what a later sub-range finds in it is deterministic;
//        pass B: every warp probes 32 positions per step: the two most recent offsets of its own parse (repeat
//                candidates), its incremental "short" table (4-byte hash, 1024 entries: the nearest earlier position
//                of its own sub-range), then the long tables of the three sub-ranges in front of it (matches across
//                sub-ranges, up to 64 KiB back; 2 tag bits per entry filter most false candidates before the 6-byte
//                check).  A ballot picks the first hit, a second ballot extends it.
//      Sequences are (literal length, match length, offset) triples, so the eight lists concatenate by adding a
//      sub-range's trailing literals to the next one's first sequence;
//   2. repeated offsets (RepeatedOffsets.java:16-49, RFC 8878 3.1.1.5): one pass over the concatenated list turns
//      offsets that equal one of the three most recent ones into repeat codes 1..3 (exact decoder history);
//   3. prefix sums over the sequences give literal-buffer offsets; literals are gathered and histogrammed;
//   4. Huffman: code lengths (depth limit 11) by one thread, then every symbol's code is OR-ed into the
//      output at its bit offset (reverse prefix sum of code lengths), 4 streams on 4 warps;
//   5. sequences (SequenceEncoder.compressSequences :66-209): code histograms, selectEncodingType (:299-341) per
//      stream -- RLE, predefined, or FSE-compressed with normalizeCounts / writeNormalizedCounts
//      (FiniteStateEntropy.java:257-521) -- three threads build the tables and walk the three state chains, then all
//      threads OR their sequences' bits at prefix-sum offsets;
//   6. block assembly, raw-block fallback when the gain is below the reference's threshold.
#include "zstd_common.cuh"
#include "zstd_fse_enc.cuh"
#include "xxh64_device.cuh"

namespace {
using namespace zs;

constexpr int kThreads = 256;
constexpr int kNQ = kThreads / 32;                         // sub-ranges of a block, one per warp
constexpr int kShortLog = 10;                              // incremental table of a sub-range: 4-byte hash
constexpr int kLongLog = 11;                               // final table of a sub-range: 6-byte hash, 14-bit position + 2 tag bits
constexpr int kPrevTables = 3;                             // long tables of this many earlier sub-ranges are probed
constexpr uint16_t kEmpty16 = 0xFFFF;
constexpr int kMaxSeqQ = kMaxBlock / kNQ / 4 + 16;       // sequences per sub-range (min match 4)
constexpr int kMaxSeq = kNQ * kMaxSeqQ;
constexpr int kStreamStage = 48 * 1024;                  // staging bytes per Huffman stream (32768 symbols x 11 bits)
constexpr int kSeqStage = kMaxBlock + 1024;              // staging bytes for the sequence bitstream
constexpr int64_t kScratchPerCta = (int64_t) kMaxSeq * 8 + (kMaxBlock + 64) + (int64_t) 3 * kMaxSeq * 2 + 4 * kStreamStage + kSeqStage + 256 +
                                   (int64_t) kMaxSeq * 8 + (int64_t) kMaxSeq * 4;   // + compact sequence list + code words

__device__ __forceinline__ uint32_t zenc_short_slot(uint32_t v) { return (v * 2654435761u) >> (32 - kShortLog); }
// 6-byte hash (hash6 of DoubleFastBlockCompressor.java:216-256): slot in the low kLongLog bits, 2 tag bits above them
__device__ __forceinline__ uint32_t zenc_long_hash(uint64_t v) { return (uint32_t) (((v << 16) * 0xCF1BBCDCBF9BULL) >> (64 - kLongLog - 2)); }

struct NodeTable { int32_t count[512]; int16_t parents[512]; int16_t symbols[512]; uint8_t nbits[512]; };

struct EncSmem {
    union {
        struct { uint16_t inc[kNQ][1 << kShortLog]; uint16_t fin[kNQ][1 << kLongLog]; } t;   // match finding
        uint32_t rep[(kNQ << kShortLog) / 2 + (kNQ << kLongLog) / 2];                           // repeat-offset pass: one word per sequence
        NodeTable nt;                                                                            // Huffman tree
    } u;
    uint32_t hist[256];
    uint16_t hcode[256];
    uint8_t hbits[256];
    uint16_t ll_next[512], ml_next[512], of_next[256];     // FSE encode tables of the current block (tableLog <= 9 / 9 / 8)
    int32_t ll_dnb[36], ll_dfs[36], ml_dnb[53], ml_dfs[53], of_dnb[32], of_dfs[32];
    int32_t chist[3][56];                                  // code histograms: OF, ML, LL
    uint8_t tdesc[3][136];                                 // table descriptions (RLE symbol or normalized counts): OF, ML, LL
    int32_t scan[kThreads + 1];
    int32_t qcount[kNQ], qtrail[kNQ], qbase[kNQ + 1];
    int32_t v[32];   // broadcast slots
    uint8_t wbuf[264];   // serialized Huffman table description (header byte + weights)
};

enum { V_NSEQ = 0, V_LASTLIT, V_NLIT, V_LITMODE, V_MAXSYM, V_HUFBITS, V_HTABLE_BYTES, V_STREAM_BYTES0, V_STREAM_BYTES1, V_STREAM_BYTES2,
       V_STREAM_BYTES3, V_SEQ_TOTAL_BITS, V_FINAL_OF, V_FINAL_ML, V_FINAL_LL, V_LOG_OF, V_LOG_ML, V_LOG_LL, V_MODE_OF, V_MODE_ML, V_MODE_LL,
       V_DESC_OF, V_DESC_ML, V_DESC_LL, V_CHECKSUM, V_REP1, V_REP2, V_REP3, V_REPT1, V_REPT2, V_REPT3 };

__device__ __forceinline__ uint64_t pack_seq(uint32_t ll, uint32_t ml, uint32_t off) { return (uint64_t) ll | ((uint64_t) ml << 20) | ((uint64_t) off << 40); }
__device__ __forceinline__ uint32_t seq_ll(uint64_t s) { return (uint32_t) (s & 0xFFFFF); }
__device__ __forceinline__ uint32_t seq_ml(uint64_t s) { return (uint32_t) ((s >> 20) & 0xFFFFF); }
__device__ __forceinline__ uint32_t seq_off(uint64_t s) { return (uint32_t) (s >> 40); }

// SequenceStore.literalLengthToCode / matchLengthToCode (zstd/SequenceStore.java:137-159) as formulas
__device__ __forceinline__ int ll_code_of(uint32_t ll)
{
    if (ll >= 64) return highbit(ll) + 19;
    if (ll < 16) return (int) ll;
    if (ll < 24) return 16 + (int) ((ll - 16) >> 1);
    if (ll < 32) return 20 + (int) ((ll - 24) >> 2);
    if (ll < 48) return 22 + (int) ((ll - 32) >> 3);
    return 24;
}
__device__ __forceinline__ int ml_code_of(uint32_t mlb)   // mlb = match length - 3
{
    if (mlb >= 128) return highbit(mlb) + 36;
    if (mlb < 32) return (int) mlb;
    if (mlb < 40) return 32 + (int) ((mlb - 32) >> 1);
    if (mlb < 48) return 36 + (int) ((mlb - 40) >> 2);
    if (mlb < 64) return 38 + (int) ((mlb - 48) >> 3);
    if (mlb < 96) return 40 + (int) ((mlb - 64) >> 4);
    return 42;
}

// FseCompressionTable.initialize (zstd/FseCompressionTable.java:52-111), single thread, tables in shared memory
__device__ void fse_build_ctable(uint16_t *next_state, int32_t *dnb, int32_t *dfs, const int16_t *norm, int max_symbol, int table_log, uint8_t *spread, int32_t *cumul)
{
    const int size = 1 << table_log;
    int high = size - 1;
    cumul[0] = 0;
    for (int i = 1; i <= max_symbol + 1; i++) {
        if (norm[i - 1] == -1) { cumul[i] = cumul[i - 1] + 1; spread[high--] = (uint8_t) (i - 1); }
        else cumul[i] = cumul[i - 1] + norm[i - 1];
    }
    cumul[max_symbol + 1] = size + 1;
    const int mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    int position = 0;
    for (int s = 0; s <= max_symbol; s++) {
        for (int i = 0; i < norm[s]; i++) {
            spread[position] = (uint8_t) s;
            do { position = (position + step) & mask; } while (position > high);
        }
    }
    for (int i = 0; i < size; i++) { int s = spread[i]; next_state[cumul[s]++] = (uint16_t) (size + i); }
    int total = 0;
    for (int s = 0; s <= max_symbol; s++) {
        int n = norm[s];
        if (n == 0) dnb[s] = ((table_log + 1) << 16) - size;
        else if (n == -1 || n == 1) { dnb[s] = (table_log << 16) - size; dfs[s] = total - 1; total++; }
        else {
            int max_bits_out = table_log - highbit((uint32_t) (n - 1));
            dnb[s] = (max_bits_out << 16) - (n << max_bits_out);
            dfs[s] = total - n;
            total += n;
        }
    }
}

__device__ __forceinline__ int fse_begin(const uint16_t *next_state, const int32_t *dnb, const int32_t *dfs, int symbol)
{
    int output_bits = (int) ((uint32_t) (dnb[symbol] + (1 << 15)) >> 16);
    int base = (int) ((uint32_t) ((output_bits << 16) - dnb[symbol]) >> output_bits);
    return next_state[base + dfs[symbol]];
}

// The leaf order of HuffmanCompressionTable.buildTree (:105-130: insertion sort, descending by count, equal counts in symbol
// order) as a rank sort over all 256 threads: thread s counts the symbols that come before symbol s.  Also clears the
// node table.  Call with all kThreads threads, __syncthreads() afterwards.
__device__ __forceinline__ void huf_sort_leaves(EncSmem &sm)
{
    NodeTable &nt = sm.u.nt;
    const int t = threadIdx.x;
    for (int i = t; i < 512; i += kThreads) { nt.count[i] = 0; nt.parents[i] = 0; nt.symbols[i] = 0; nt.nbits[i] = 0; }
    __syncthreads();
    if (t < 256) {
        const uint32_t mine = sm.hist[t];
        int rank = 0;
        for (int s = 0; s < 256; s++) { const uint32_t c = sm.hist[s]; rank += (c > mine) || (c == mine && s < t); }
        nt.count[rank] = (int32_t) mine;
        nt.symbols[rank] = (int16_t) t;
    }
}

// HuffmanCompressionTable.buildTree + setMaxHeight (zstd/HuffmanCompressionTable.java:105-190, :294-390), single thread, on
// the leaves huf_sort_leaves ordered.  Produces code lengths in hbits[] and canonical values in hcode[]; returns the
// maximum code length used.
__device__ int huf_build(EncSmem &sm, int max_symbol, int max_bits)
{
    NodeTable &nt = sm.u.nt;   // leaves sorted by huf_sort_leaves, the rest zeroed
    int current = 0;
    int last_non_zero = max_symbol;
    while (nt.count[last_non_zero] == 0) last_non_zero--;
    const int non_leaf_start = 256;
    current = non_leaf_start;
    int current_leaf = last_non_zero, current_non_leaf = current;
    nt.count[current] = nt.count[current_leaf] + nt.count[current_leaf - 1];
    nt.parents[current_leaf] = (int16_t) current;
    nt.parents[current_leaf - 1] = (int16_t) current;
    current++;
    current_leaf -= 2;
    const int root = 256 + last_non_zero - 1;
    for (int n = current; n <= root; n++) nt.count[n] = 1 << 30;
    while (current <= root) {
        int child1, child2;
        if (current_leaf >= 0 && nt.count[current_leaf] < nt.count[current_non_leaf]) child1 = current_leaf--; else child1 = current_non_leaf++;
        if (current_leaf >= 0 && nt.count[current_leaf] < nt.count[current_non_leaf]) child2 = current_leaf--; else child2 = current_non_leaf++;
        nt.count[current] = nt.count[child1] + nt.count[child2];
        nt.parents[child1] = (int16_t) current;
        nt.parents[child2] = (int16_t) current;
        current++;
    }
    nt.nbits[root] = 0;
    for (int n = root - 1; n >= non_leaf_start; n--) nt.nbits[n] = (uint8_t) (nt.nbits[nt.parents[n]] + 1);
    for (int n = 0; n <= last_non_zero; n++) nt.nbits[n] = (uint8_t) (nt.nbits[nt.parents[n]] + 1);

    // depth limit (setMaxHeight)
    int largest_bits = nt.nbits[last_non_zero];
    if (largest_bits > max_bits) {
        int total_cost = 0;
        const int base_cost = 1 << (largest_bits - max_bits);
        int n = last_non_zero;
        while (nt.nbits[n] > max_bits) {
            total_cost += base_cost - (1 << (largest_bits - nt.nbits[n]));
            nt.nbits[n] = (uint8_t) max_bits;
            n--;
        }
        while (nt.nbits[n] == max_bits) n--;
        total_cost = (int) ((uint32_t) total_cost >> (largest_bits - max_bits));
        const int no_symbol = (int) 0xF0F0F0F0;
        int rank_last[14];
        for (int i = 0; i < 14; i++) rank_last[i] = no_symbol;
        int current_bits = max_bits;
        for (int pos = n; pos >= 0; pos--) {
            if (nt.nbits[pos] >= current_bits) continue;
            current_bits = nt.nbits[pos];
            rank_last[max_bits - current_bits] = pos;
        }
        while (total_cost > 0) {
            int dec = highbit((uint32_t) total_cost) + 1;
            for (; dec > 1; dec--) {
                int high_pos = rank_last[dec], low_pos = rank_last[dec - 1];
                if (high_pos == no_symbol) continue;
                if (low_pos == no_symbol) break;
                if (nt.count[high_pos] <= 2 * nt.count[low_pos]) break;
            }
            while (dec <= 12 && rank_last[dec] == no_symbol) dec++;
            total_cost -= 1 << (dec - 1);
            if (rank_last[dec - 1] == no_symbol) rank_last[dec - 1] = rank_last[dec];
            nt.nbits[rank_last[dec]]++;
            if (rank_last[dec] == 0) rank_last[dec] = no_symbol;
            else {
                rank_last[dec]--;
                if (nt.nbits[rank_last[dec]] != max_bits - dec) rank_last[dec] = no_symbol;
            }
        }
        while (total_cost < 0) {
            if (rank_last[1] == no_symbol) {
                while (nt.nbits[n] == max_bits) n--;
                nt.nbits[n + 1]--;
                rank_last[1] = n + 1;
                total_cost++;
                continue;
            }
            nt.nbits[rank_last[1] + 1]--;
            rank_last[1]++;
            total_cost++;
        }
        largest_bits = max_bits;
    }
    // canonical codes (HuffmanCompressionTable.initialize :80-100)
    for (int s = 0; s < 256; s++) sm.hbits[s] = 0;
    for (int node = 0; node <= last_non_zero; node++) sm.hbits[nt.symbols[node]] = nt.nbits[node];
    int entries[13], values[13];
    for (int i = 0; i < 13; i++) { entries[i] = 0; values[i] = 0; }
    for (int n = 0; n <= last_non_zero; n++) entries[nt.nbits[n]]++;
    int starting = 0;
    for (int rank = largest_bits; rank > 0; rank--) {
        values[rank] = starting;
        starting = (starting + entries[rank]) >> 1;
    }
    for (int s = 0; s <= max_symbol; s++) sm.hcode[s] = sm.hbits[s] ? (uint16_t) values[sm.hbits[s]]++ : 0;
    return largest_bits;
}

// SequenceEncoder.selectEncodingType (zstd/SequenceEncoder.java:299-341) for strategy DFAST (ordinal 1): 0 basic (predefined
// table), 1 RLE, 2 FSE-compressed.
__device__ __forceinline__ int select_seq_encoding(int largest, int nseq, int default_log, bool default_allowed)
{
    if (largest == nseq) return (default_allowed && nseq <= 2) ? 0 : 1;
    if (default_allowed) {
        const int min_sequences = ((1 << default_log) * 9) >> 3;
        if (nseq < min_sequences || largest < (nseq >> (default_log - 1))) return 0;
    }
    return 2;
}

// One sequence stream (k: 0 offsets, 1 match lengths, 2 literal lengths) of SequenceEncoder.compressSequences (:96-199):
// picks the encoding from the code histogram, builds the FSE encode table into shared memory and serialises its
// description (nothing for the predefined table, the symbol for RLE, writeNormalizedCounts for a compressed table).
// Single thread.  last_code = code of the last sequence (buildCompressionTable :211-226 leaves it out of the statistics).
__device__ void build_seq_table(EncSmem &sm, int k, int nseq, int last_code)
{
    uint16_t *nx = k == 0 ? sm.of_next : k == 1 ? sm.ml_next : sm.ll_next;
    int32_t *dnb = k == 0 ? sm.of_dnb : k == 1 ? sm.ml_dnb : sm.ll_dnb;
    int32_t *dfs = k == 0 ? sm.of_dfs : k == 1 ? sm.ml_dfs : sm.ll_dfs;
    int32_t *counts = sm.chist[k];
    const int16_t *def_norm = k == 0 ? kDefOF : k == 1 ? kDefML : kDefLL;
    const int def_max = k == 0 ? 28 : k == 1 ? 52 : 35, def_log = k == 0 ? 5 : 6, max_log = k == 0 ? 8 : 9;
    int max_symbol = k == 0 ? 31 : def_max;
    while (counts[max_symbol] == 0) max_symbol--;
    int largest = 0;
    for (int i = 0; i <= max_symbol; i++) largest = max(largest, counts[i]);
    const bool default_allowed = k != 0 || max_symbol < 28;        // :141
    const int mode = select_seq_encoding(largest, nseq, def_log, default_allowed);
    uint8_t spread[512];
    int32_t cumul[56];
    int table_log = def_log, desc = 0;
    if (mode == 1) {
        // FseCompressionTable.initializeRleTable (:41-50): one symbol, zero bits per step
        sm.tdesc[k][0] = (uint8_t) max_symbol;
        desc = 1;
        table_log = 0;
        nx[0] = 0; nx[1] = 0;
        dnb[max_symbol] = 0; dfs[max_symbol] = 0;
    }
    else if (mode == 0) fse_build_ctable(nx, dnb, dfs, def_norm, def_max, def_log, spread, cumul);
    else {
        int16_t norm[56];
        table_log = fse_optimal_table_log(max_log, nseq, max_symbol);
        int total = nseq;
        if (counts[last_code] > 1) { counts[last_code]--; total--; }
        fse_normalize(norm, table_log, counts, total, max_symbol);
        fse_build_ctable(nx, dnb, dfs, norm, max_symbol, table_log, spread, cumul);
        desc = fse_write_ncount(sm.tdesc[k], (int) sizeof(sm.tdesc[k]), norm, max_symbol, table_log);
    }
    sm.v[V_MODE_OF + k] = mode;
    sm.v[V_LOG_OF + k] = table_log;
    sm.v[V_DESC_OF + k] = desc;
}

// OR `nbits` bits of `value` into a zero-initialised little-endian bit buffer at bit position `pos`
__device__ __forceinline__ void or_bits(uint32_t *buf, uint32_t pos, uint64_t value, int nbits)
{
    if (nbits == 0) return;
    uint32_t w = pos >> 5, sh = pos & 31;
    uint64_t lo = value << sh;
    atomicOr(buf + w, (uint32_t) lo);
    if (sh + nbits > 32) atomicOr(buf + w + 1, (uint32_t) (lo >> 32));
    if (sh + nbits > 64) atomicOr(buf + w + 2, (uint32_t) (value >> (64 - sh)));
}

// exclusive block scan of one int per thread (kThreads threads); returns the exclusive prefix, *total gets the sum
__device__ __forceinline__ int block_scan_excl(EncSmem &sm, int v, int *total)
{
    const int tid = threadIdx.x;
    __syncthreads();
    sm.scan[tid] = v;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < kThreads; i++) { int t = sm.scan[i]; sm.scan[i] = acc; acc += t; }
        sm.scan[kThreads] = acc;
    }
    __syncthreads();
    *total = sm.scan[kThreads];
    return sm.scan[tid];
}

template <bool kCg = false>
__device__ __forceinline__ void block_copy(uint8_t *dst, const uint8_t *src, int64_t n)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t per = ((n + kNQ - 1) / kNQ + 15) & ~15LL;
    int64_t a = per * warp, b = a + per < n ? a + per : n;
    if (a < b) warp_copy<kCg>(dst + a, src + a, b - a, lane);
}

__global__ void __launch_bounds__(kThreads) zstd_compress_kernel(AccBatch b, uint8_t *scratch_base, const int64_t *frame_hashes)
{
    extern __shared__ __align__(16) uint8_t zenc_smem[];
    EncSmem &sm = *reinterpret_cast<EncSmem *>(zenc_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *scratch = scratch_base + (int64_t) blockIdx.x * kScratchPerCta;
    uint64_t *seqs = reinterpret_cast<uint64_t *>(scratch);                              // [4][kMaxSeqQ]
    uint8_t *lits = scratch + (int64_t) kMaxSeq * 8;                                     // gathered literals
    uint16_t *sbits = reinterpret_cast<uint16_t *>(lits + kMaxBlock + 64);               // [3][kMaxSeq] state bits of OF, ML, LL
    uint32_t *stage = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(sbits) + (int64_t) 3 * kMaxSeq * 2);   // 4 stream stages + seq stage
    uint32_t *seq_stage = stage + 4 * kStreamStage / 4;
    uint64_t *seqc = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(seq_stage) + kSeqStage + 256);   // concatenated list, offsets as offset values
    uint32_t *codew_g = reinterpret_cast<uint32_t *>(seqc + kMaxSeq);                                            // OF | ML << 8 | LL << 16 codes per sequence

    for (;;) {
        __syncthreads();
        if (tid == 0) sm.v[31] = (int32_t) atomicAdd(b.work_counter, 1u);
        __syncthreads();
        const unsigned int idx = (unsigned int) sm.v[31];
        if ((int64_t) idx >= b.n) break;
        const uint8_t *in = b.src + b.src_off[idx];
        const int64_t in_len = b.src_len[idx];
        uint8_t *out = b.dst + b.dst_off[idx];
        const int64_t out_cap = b.dst_cap[idx];
        int64_t bound = in_len + (in_len >> 8) + (in_len < kMaxBlock ? ((kMaxBlock - in_len) >> 11) : 0);
        if (in_len > 0x7fffffff || out_cap < bound) {
            if (tid == 0) { b.out_len[idx] = 0; b.status[idx] = ACC_STATUS(ACC_E_ARGUMENT, ACC_R_MAX_OUTPUT_TOO_SMALL); }
            continue;
        }
        // ---- frame header (ZstdFrameCompressor.writeMagic / writeFrameHeader :52-120) ----
        int64_t op = 0;
        {
            const bool single_segment = in_len <= kMaxBlock;   // window = max(input, 128 KiB): blocks never reference beyond themselves
            const int cs_desc = (in_len >= 256) + (in_len >= 65536 + 256);
            if (tid == 0) {
                out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
                out[4] = (uint8_t) ((cs_desc << 6) | 0x4 | (single_segment ? 0x20 : 0));
                int p = 5;
                if (!single_segment) out[p++] = (uint8_t) ((17 - 10) << 3);   // window 2^17
                if (cs_desc == 0) { if (single_segment) out[p++] = (uint8_t) in_len; }
                else if (cs_desc == 1) { uint32_t v = (uint32_t) (in_len - 256); out[p++] = (uint8_t) v; out[p++] = (uint8_t) (v >> 8); }
                else { uint32_t v = (uint32_t) in_len; for (int k = 0; k < 4; k++) out[p++] = (uint8_t) (v >> (8 * k)); }
            }
            op = 5 + (single_segment ? 0 : 1) + (cs_desc == 0 ? (single_segment ? 1 : 0) : (cs_desc == 1 ? 2 : 4));
        }
        // frame checksum: XXH64 of the whole input, computed for the batch by xxh64_kernel before this kernel
        if (tid == 0) {
            sm.v[V_CHECKSUM] = (int32_t) (uint32_t) (uint64_t) frame_hashes[idx];
            sm.v[V_REP1] = 1; sm.v[V_REP2] = 4; sm.v[V_REP3] = 8;   // repeated offsets at the start of a frame (RepeatedOffsets.java:18-19 + RFC 8878)
        }

        int64_t block_start = 0;
        bool last_block;
        do {
            const int64_t remaining = in_len - block_start;
            const int bl = (int) (remaining < kMaxBlock ? remaining : kMaxBlock);
            last_block = remaining <= kMaxBlock;
            const uint8_t *blk = in + block_start;
            bool compressed = false;
            int64_t block_bytes = 0;   // payload bytes after the 3-byte block header

            if (bl >= 16) {
                // ================= 1. match finding =================
                {
                    uint16_t *tz = &sm.u.t.inc[0][0];      // inc and fin are contiguous
                    for (int i = tid; i < (kNQ << kShortLog) + (kNQ << kLongLog); i += kThreads) tz[i] = kEmpty16;
                }
                __syncthreads();
                {
                    const int q = ((bl + kNQ - 1) / kNQ + 31) & ~31;        // <= 16384: positions inside a sub-range fit 14 bits
                    const int qs = min(q * warp, bl), qe = min(qs + q, bl);
                    // ---- pass A: the long table of this sub-range (last position of every 6-byte value) ----
                    {
                        uint16_t *fin = sm.u.t.fin[warp];
                        const int lim = min(qe, bl - 8);                   // 8 readable bytes at every hashed position
                        for (int base = qs; base < lim; base += 32) {
                            const int p = base + lane;
                            uint32_t h = 0, slot = 0xFFFFFFFFu - (uint32_t) lane;   // idle lanes: distinct dummies
                            if (p < lim) { h = zenc_long_hash(ld_u64_unaligned(blk + p)); slot = h & ((1u << kLongLog) - 1); }
                            if (p < lim) fin[slot] = (uint16_t) ((uint32_t) (p - qs) | ((h >> kLongLog) << 14));   // equal slots in one step: see lz4.cu
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                    // ---- pass B: greedy parse of the sub-range, 32 positions per step ----
                    uint16_t *inc = sm.u.t.inc[warp];
                    uint64_t *my = seqs + (int64_t) warp * kMaxSeqQ;
                    int count = 0, anchor = qs, pos = qs;
                    int rep1 = 0, rep2 = 0;                                  // the two most recent offsets of this parse (0: none yet)
                    const int probe_limit = min(qe - 3, bl - 8);             // a match (>= 4 bytes) ends inside its sub-range
                    while (pos < probe_limit) {
                        const int p = pos + lane;
                        const bool live = p < probe_limit;
                        int cand = -1;
                        uint32_t sslot = 0xFFFFFFFFu - (uint32_t) lane;
                        if (live) {
                            // every candidate's bytes are requested before the first comparison: the loads overlap instead of
                            // forming a chain of dependent round trips
                            const uint8_t *pp = blk + p;
                            const uint64_t cur8 = ld_u64_unaligned(pp);
                            const bool t1 = rep1 && p >= rep1, t2 = rep2 && p >= rep2;
                            const uint32_t v1 = t1 ? ld_u32_unaligned(pp - rep1) : 0u, v2 = t2 ? ld_u32_unaligned(pp - rep2) : 0u;
                            const uint32_t cur = (uint32_t) cur8;
                            sslot = zenc_short_slot(cur);
                            const uint32_t e = inc[sslot];
                            const bool ti = e != kEmpty16 && qs + (int) e < p;
                            const uint32_t vi = ti ? ld_u32_unaligned(blk + (qs + (int) e)) : 0u;
                            const uint32_t h = zenc_long_hash(cur8);
                            const uint32_t ls = h & ((1u << kLongLog) - 1), tag = h >> kLongLog;
                            int cl[kPrevTables];
                            uint64_t wl[kPrevTables];
#pragma unroll
                            for (int k = 1; k <= kPrevTables; k++) {
                                cl[k - 1] = -1; wl[k - 1] = 0;
                                if (k <= warp) {
                                    const uint32_t f = sm.u.t.fin[warp - k][ls];
                                    if (f != kEmpty16 && (f >> 14) == tag) {
                                        cl[k - 1] = q * (warp - k) + (int) (f & 0x3FFFu);
                                        wl[k - 1] = ld_u64_unaligned(blk + cl[k - 1]);
                                    }
                                }
                            }
                            if (t1 && v1 == cur) cand = p - rep1;
                            else if (t2 && v2 == cur) cand = p - rep2;
                            else if (ti && vi == cur) cand = qs + (int) e;
                            else {
#pragma unroll
                                for (int k = 0; k < kPrevTables; k++)
                                    if (cand < 0 && cl[k] >= 0 && ((wl[k] ^ cur8) << 16) == 0) cand = cl[k];
                            }
                        }
                        const unsigned hits = __ballot_sync(kFull, cand >= 0);
                        const int first_hit = hits ? __ffs(hits) - 1 : 31;
                        // insert after the lookups, and only positions up to the first match (see lz4.cu)
                        if (live && lane <= first_hit) inc[sslot] = (uint16_t) (p - qs);
                        __syncwarp();
                        if (hits == 0) { pos += 32; continue; }
                        int mpos = pos + first_hit;
                        int ref = __shfl_sync(kFull, cand, first_hit);
                        while (mpos > anchor && ref > 0 && blk[mpos - 1] == blk[ref - 1]) { --mpos; --ref; }
                        int mlen = 4;
                        for (;;) {
                            const int qq = mpos + mlen + lane;
                            const bool same = qq < qe && blk[qq] == blk[ref + mlen + lane];
                            const unsigned eq = __ballot_sync(kFull, same);
                            if (eq == kFull) { mlen += 32; continue; }
                            mlen += __ffs(~eq) - 1;
                            break;
                        }
                        const int off = mpos - ref;
                        if (lane == 0) my[count] = pack_seq((uint32_t) (mpos - anchor), (uint32_t) mlen, (uint32_t) off);
                        count++;
                        if (off != rep1) { rep2 = rep1; rep1 = off; }
                        pos = mpos + mlen;
                        anchor = pos;
                    }
                    if (lane == 0) { sm.qcount[warp] = count; sm.qtrail[warp] = qe - anchor; }
                }
                __syncthreads();
                // ================= 2. concatenate the quarter lists, gather literals =================
                if (tid == 0) {
                    int carry = 0, base = 0;
                    for (int w = 0; w < kNQ; w++) {
                        sm.qbase[w] = base;
                        if (sm.qcount[w] > 0) {
                            uint64_t *first = seqs + (int64_t) w * kMaxSeqQ;
                            uint64_t s = *first;
                            *first = pack_seq(seq_ll(s) + (uint32_t) carry, seq_ml(s), seq_off(s));
                            carry = sm.qtrail[w];
                        }
                        else carry += sm.qtrail[w];
                        base += sm.qcount[w];
                    }
                    sm.qbase[kNQ] = base;
                    sm.v[V_NSEQ] = base;
                    sm.v[V_LASTLIT] = carry;
                }
                for (int i = tid; i < 256; i += kThreads) sm.hist[i] = 0;
                __syncthreads();
                const int nseq = sm.v[V_NSEQ], last_lit = sm.v[V_LASTLIT];
                auto seq_raw = [&](int i) -> uint64_t {       // sequence i of the eight per-warp lists
                    int w = 0;
#pragma unroll
                    for (int k = 1; k < kNQ; k++) w += (i >= sm.qbase[k]);
                    return seqs[(int64_t) w * kMaxSeqQ + (i - sm.qbase[w])];
                };
                auto seq_at = [&](int i) -> uint64_t { return seqc[i]; };   // the concatenated list (offsets become offset values in step 3b)
                // the eight lists as one contiguous array: every later stage indexes it directly
                for (int i = tid; i < nseq; i += kThreads) seqc[i] = seq_raw(i);
                __syncthreads();
                const int chunk = (nseq + kThreads - 1) / kThreads;
                const int c0 = min(tid * chunk, nseq), c1 = min(c0 + chunk, nseq);
                int my_ll = 0, my_all = 0;
                for (int i = c0; i < c1; i++) { uint64_t s = seq_at(i); my_ll += (int) seq_ll(s); my_all += (int) (seq_ll(s) + seq_ml(s)); }
                int tot_ll, tot_all;
                int lo = block_scan_excl(sm, my_ll, &tot_ll);
                int po = block_scan_excl(sm, my_all, &tot_all);
                const int nlit = tot_ll + last_lit;
                for (int i = c0; i < c1; i++) {
                    uint64_t s = seq_at(i);
                    const int ll = (int) seq_ll(s);
                    for (int k = 0; k < ll; k++) { uint8_t c = blk[po + k]; lits[lo + k] = c; atomicAdd(&sm.hist[c], 1u); }
                    lo += ll;
                    po += ll + (int) seq_ml(s);
                }
                for (int k = tid; k < last_lit; k += kThreads) { uint8_t c = blk[bl - last_lit + k]; lits[tot_ll + k] = c; atomicAdd(&sm.hist[c], 1u); }
                __syncthreads();

                // ================= 3. literals section planning (ZstdFrameCompressor.encodeLiterals :262-378) =================
                huf_sort_leaves(sm);
                __syncthreads();
                if (warp == kNQ - 1) {
                    // ================= 3b. repeated offsets (RFC 8878 3.1.1.5, RepeatedOffsets.java:16-49), concurrently with 3 =====
                    // Lane 0 walks the list with the decoder's history and turns every offset into its offset VALUE: 1..3 for a
                    // repeat code, offset + 3 otherwise.  The list passes through shared memory in tiles (behind the Huffman node
                    // table thread 0 is using meanwhile).
                    constexpr int kTileAt = 2048;                                   // words; sizeof(NodeTable) < 8 KiB
                    constexpr int kTile = (int) (sizeof(sm.u.rep) / 4) - kTileAt;
                    static_assert(sizeof(NodeTable) <= kTileAt * 4, "node table overlaps the repeat-offset tile");
                    uint32_t *tile = sm.u.rep + kTileAt;
                    int r1 = sm.v[V_REP1], r2 = sm.v[V_REP2], r3 = sm.v[V_REP3];      // lane 0's copy is the one that counts
                    for (int t0 = 0; t0 < nseq; t0 += kTile) {
                        const int tn = min(kTile, nseq - t0);
                        for (int i = lane; i < tn; i += 32) { const uint64_t sq = seqc[t0 + i]; tile[i] = seq_off(sq) | (seq_ll(sq) ? 0x80000000u : 0u); }
                        __syncwarp();
                        if (lane == 0) {
                            for (int i = 0; i < tn; i++) {
                                const uint32_t e = tile[i];
                                const int o = (int) (e & 0x7FFFFFFFu);
                                uint32_t val = (uint32_t) o + 3;
                                if (e >> 31) {                      // literals in front of the match
                                    if (o == r1) val = 1;
                                    else if (o == r2) { val = 2; r2 = r1; r1 = o; }
                                    else { if (o == r3) val = 3; r3 = r2; r2 = r1; r1 = o; }
                                }
                                else {                              // no literals: the codes mean rep2, rep3, rep1 - 1
                                    if (o == r2) { val = 1; r2 = r1; r1 = o; }
                                    else { if (o == r3) val = 2; else if (o == r1 - 1) val = 3; r3 = r2; r2 = r1; r1 = o; }
                                }
                                tile[i] = val;
                            }
                        }
                        __syncwarp();
                        for (int i = lane; i < tn; i += 32) {
                            const uint64_t sq = seqc[t0 + i];
                            seqc[t0 + i] = pack_seq(seq_ll(sq), seq_ml(sq), tile[i]);
                        }
                        __syncwarp();
                    }
                    if (lane == 0) { sm.v[V_REPT1] = r1; sm.v[V_REPT2] = r2; sm.v[V_REPT3] = r3; }
                }
                if (tid == 0) {
                    int mode = 0;   // 0 raw, 1 rle, 2 huffman
                    int max_symbol = 255, largest = 0, hbits = 0, table_bytes = 0;
                    if (nlit > 63) {
                        while (sm.hist[max_symbol] == 0) max_symbol--;
                        for (int s = 0; s <= max_symbol; s++) largest = max(largest, (int) sm.hist[s]);
                        if (largest == nlit) mode = 1;
                        else if (largest <= (nlit >> 7) + 4) mode = 0;
                        else {
                            // optimalNumberOfBits(11, nlit, maxSymbol) (HuffmanCompressionTable.java:41-58)
                            int r = 11, v1 = highbit((uint32_t) (nlit - 1)) - 1;
                            if (v1 < r) r = v1;
                            int mtl = min(highbit((uint32_t) (nlit - 1)) + 1, highbit((uint32_t) max_symbol) + 2);
                            if (mtl > r) r = mtl;
                            if (r < 5) r = 5;
                            if (r > 12) r = 12;
                            hbits = huf_build(sm, max_symbol, r);
                            // table description (HuffmanCompressionTable.write :202-263): FSE-compressed weights when that is
                            // smaller than the direct 4-bit form (and mandatory above 128 entries, which the direct form cannot hold)
                            uint8_t weights[256];
                            for (int s2 = 0; s2 < max_symbol; s2++) weights[s2] = sm.hbits[s2] ? (uint8_t) (hbits + 1 - sm.hbits[s2]) : 0;
                            const int fsz = fse_compress_weights(sm.wbuf + 1, 250, weights, max_symbol);
                            if (fsz > 1 && fsz < max_symbol / 2 && fsz < 128) {
                                sm.wbuf[0] = (uint8_t) fsz;
                                table_bytes = 1 + fsz;
                                mode = 2;
                            }
                            else if (max_symbol <= 128) {
                                sm.wbuf[0] = (uint8_t) (127 + max_symbol);
                                weights[max_symbol] = 0;
                                for (int i = 0; i < max_symbol; i += 2) sm.wbuf[1 + i / 2] = (uint8_t) ((weights[i] << 4) + weights[i + 1]);
                                table_bytes = 1 + (max_symbol + 1) / 2;
                                mode = 2;
                            }
                            else mode = 0;
                        }
                    }
                    sm.v[V_NLIT] = nlit; sm.v[V_LITMODE] = mode; sm.v[V_MAXSYM] = max_symbol; sm.v[V_HUFBITS] = hbits; sm.v[V_HTABLE_BYTES] = table_bytes;
                }
                __syncthreads();
                int lit_mode = sm.v[V_LITMODE];
                const bool single_stream = nlit < 256;
                const int seg = (nlit + 3) / 4;
                int lit_section = 0, lit_header = 0;
                if (lit_mode == 2) {
                    // stream sizes: every warp sums the code lengths of its stream (warp 0 alone for a single stream)
                    const int nstreams = single_stream ? 1 : 4;
                    if (warp < nstreams) {
                        const int a = single_stream ? 0 : seg * warp, e = single_stream ? nlit : min(a + seg, nlit);
                        int bits = 0;
                        for (int i = a + lane; i < e; i += 32) bits += sm.hbits[lits[i]];
                        for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(kFull, bits, o);
                        if (lane == 0) sm.v[V_STREAM_BYTES0 + warp] = (bits + 1 + 7) >> 3;   // + end mark
                    }
                    __syncthreads();
                    int streams_total = single_stream ? sm.v[V_STREAM_BYTES0]
                                                      : 6 + sm.v[V_STREAM_BYTES0] + sm.v[V_STREAM_BYTES1] + sm.v[V_STREAM_BYTES2] + sm.v[V_STREAM_BYTES3];
                    const int total = sm.v[V_HTABLE_BYTES] + streams_total;
                    bool ok = total < nlit - ((nlit >> 6) + 2);
                    if (!single_stream) ok = ok && sm.v[V_STREAM_BYTES0] < 65536 && sm.v[V_STREAM_BYTES1] < 65536 && sm.v[V_STREAM_BYTES2] < 65536;
                    if (!ok) lit_mode = 0;
                    else { lit_header = 3 + (nlit >= 1024) + (nlit >= 16384); lit_section = lit_header + total; }
                }
                if (lit_mode == 0) { lit_header = 1 + (nlit >= 32) + (nlit >= 4096); lit_section = lit_header + nlit; }
                else if (lit_mode == 1) { lit_header = 1 + (nlit > 31) + (nlit > 4095); lit_section = lit_header + 1; }

                // ================= 4. sequences section planning =================
                int seq_header = 1;       // nseq == 0: just the count byte
                int seq_bytes = 0;
                if (nseq > 0) {
                    // code histograms of the three streams
                    for (int i = tid; i < 3 * 56; i += kThreads) (&sm.chist[0][0])[i] = 0;
                    __syncthreads();
                    // one word of codes per sequence, in shared memory when the list fits (the chain walkers below are serial:
                    // their loads must be short)
                    uint32_t *codew = nseq <= (int) (sizeof(sm.u.rep) / 4) ? sm.u.rep : codew_g;
                    for (int i = tid; i < nseq; i += kThreads) {
                        const uint64_t s = seq_at(i);
                        const int ofc = highbit(seq_off(s)), mlc = ml_code_of(seq_ml(s) - 3), llc = ll_code_of(seq_ll(s));
                        codew[i] = (uint32_t) ofc | ((uint32_t) mlc << 8) | ((uint32_t) llc << 16);
                        atomicAdd(&sm.chist[0][ofc], 1);
                        atomicAdd(&sm.chist[1][mlc], 1);
                        atomicAdd(&sm.chist[2][llc], 1);
                    }
                    __syncthreads();
                    // three threads: table selection + construction, then the FSE state chain of their stream from the last
                    // sequence to the first (SequenceEncoder.encodeSequences :228-297)
                    if (tid == 0 || tid == 32 || tid == 64) {
                        const int k = tid >> 5;   // 0 OF, 1 ML, 2 LL
                        auto code_at = [&](int i) { return (int) ((codew[i] >> (8 * k)) & 0xFF); };
                        const int last_code = code_at(nseq - 1);
                        build_seq_table(sm, k, nseq, last_code);
                        const uint16_t *nx = k == 0 ? sm.of_next : k == 1 ? sm.ml_next : sm.ll_next;
                        const int32_t *dnb = k == 0 ? sm.of_dnb : k == 1 ? sm.ml_dnb : sm.ll_dnb;
                        const int32_t *dfs = k == 0 ? sm.of_dfs : k == 1 ? sm.ml_dfs : sm.ll_dfs;
                        uint16_t *sb = sbits + (int64_t) k * kMaxSeq;
                        int state = sm.v[V_MODE_OF + k] == 1 ? 0 : fse_begin(nx, dnb, dfs, last_code);
                        sb[nseq - 1] = 0;
                        // the state is the only loop-carried value: the code of the next step and its two table entries are
                        // loaded one step ahead, so a step waits for ONE dependent shared-memory load (next_state)
                        int code = nseq >= 2 ? code_at(nseq - 2) : 0;
                        int d_nb = dnb[code], d_fs = dfs[code];
                        for (int i = nseq - 2; i >= 0; i--) {
                            const int code2 = i > 0 ? code_at(i - 1) : code;
                            const int d_nb2 = dnb[code2], d_fs2 = dfs[code2];
                            const int nb = (int) ((uint32_t) (state + d_nb) >> 16);
                            sb[i] = (uint16_t) ((state & ((1 << nb) - 1)) | (nb << 12));
                            state = nx[(state >> nb) + d_fs];
                            d_nb = d_nb2; d_fs = d_fs2;
                        }
                        sm.v[V_FINAL_OF + k] = state;
                    }
                    __syncthreads();
                    const int log_of = sm.v[V_LOG_OF], log_ml = sm.v[V_LOG_ML], log_ll = sm.v[V_LOG_LL];
                    seq_header = (nseq < 0x7F ? 1 : (nseq < 0x7F00 ? 2 : 3)) + 1 + sm.v[V_DESC_LL] + sm.v[V_DESC_OF] + sm.v[V_DESC_ML];
                    // bits per sequence, reverse prefix sums (encode order is last -> first)
                    int my_bits = 0;
                    for (int i = c0; i < c1; i++) {
                        const uint32_t cw = codew[i];
                        const int ofc = (int) (cw & 0xFF), mlc = (int) ((cw >> 8) & 0xFF), llc = (int) (cw >> 16);
                        my_bits += kLLBits[llc] + kMLBits[mlc] + ofc + (sbits[i] >> 12) + (sbits[kMaxSeq + i] >> 12) + (sbits[2 * kMaxSeq + i] >> 12);
                    }
                    int total_bits;
                    const int before = block_scan_excl(sm, my_bits, &total_bits);
                    const int my_start = total_bits - before - my_bits;   // bits emitted before this thread's chunk in encode order
                    if (tid == 0) sm.v[V_SEQ_TOTAL_BITS] = total_bits;
                    seq_bytes = (total_bits + log_ml + log_of + log_ll + 1 + 7) >> 3;
                    const int seq_section_try = seq_header + seq_bytes;
                    // ================= 5. decide, then emit =================
                    const int payload = lit_section + seq_section_try;
                    compressed = payload <= bl - ((bl >> 6) + 2) && seq_bytes < kSeqStage - 64;
                    if (compressed) {
                        for (int i = tid; i < (seq_bytes + 8) / 4 + 1; i += kThreads) seq_stage[i] = 0;
                        __syncthreads();
                        uint32_t bitpos = (uint32_t) my_start;
                        for (int i = c1 - 1; i >= c0; i--) {
                            const uint64_t s = seq_at(i);
                            const uint32_t ll = seq_ll(s), mlb = seq_ml(s) - 3, ofv = seq_off(s);
                            const uint32_t cw = codew[i];
                            const int ofc = (int) (cw & 0xFF), mlc = (int) ((cw >> 8) & 0xFF), llc = (int) (cw >> 16);
                            const int llb = kLLBits[llc], mlbits = kMLBits[mlc];
                            const uint32_t s_of = sbits[i], s_ml = sbits[kMaxSeq + i], s_ll = sbits[2 * kMaxSeq + i];
                            or_bits(seq_stage, bitpos, s_of & 0xFFF, (int) (s_of >> 12)); bitpos += s_of >> 12;
                            or_bits(seq_stage, bitpos, s_ml & 0xFFF, (int) (s_ml >> 12)); bitpos += s_ml >> 12;
                            or_bits(seq_stage, bitpos, s_ll & 0xFFF, (int) (s_ll >> 12)); bitpos += s_ll >> 12;
                            or_bits(seq_stage, bitpos, ll & ((1u << llb) - 1), llb); bitpos += llb;
                            or_bits(seq_stage, bitpos, mlb & ((1u << mlbits) - 1), mlbits); bitpos += mlbits;
                            or_bits(seq_stage, bitpos, ofv & ((1u << ofc) - 1), ofc); bitpos += ofc;
                        }
                        if (tid == 0) {
                            uint32_t bp = (uint32_t) total_bits;
                            or_bits(seq_stage, bp, (uint32_t) sm.v[V_FINAL_ML] & ((1u << log_ml) - 1), log_ml); bp += log_ml;
                            or_bits(seq_stage, bp, (uint32_t) sm.v[V_FINAL_OF] & ((1u << log_of) - 1), log_of); bp += log_of;
                            or_bits(seq_stage, bp, (uint32_t) sm.v[V_FINAL_LL] & ((1u << log_ll) - 1), log_ll); bp += log_ll;
                            or_bits(seq_stage, bp, 1, 1);
                        }
                    }
                }
                else {
                    compressed = lit_section + seq_header <= bl - ((bl >> 6) + 2);   // no sequences: literals + the count byte
                }
                __syncthreads();
                if (compressed) {
                    uint8_t *o = out + op + 3;
                    // ---- literals section ----
                    if (lit_mode == 2) {
                        const int nstreams = single_stream ? 1 : 4;
                        const int sb0 = sm.v[V_STREAM_BYTES0], sb1 = sm.v[V_STREAM_BYTES1], sb2 = sm.v[V_STREAM_BYTES2], sb3 = sm.v[V_STREAM_BYTES3];
                        const int streams_total = single_stream ? sb0 : 6 + sb0 + sb1 + sb2 + sb3;
                        const int total = sm.v[V_HTABLE_BYTES] + streams_total;
                        for (int i = tid; i < 4 * kStreamStage / 4; i += kThreads) {
                            // zero only what will be used
                            const int st = i / (kStreamStage / 4), wd = i % (kStreamStage / 4);
                            const int need = st < nstreams ? (sm.v[V_STREAM_BYTES0 + st] + 8) / 4 + 1 : 0;
                            if (wd < need) stage[i] = 0;
                        }
                        __syncthreads();
                        if (warp < nstreams) {
                            const int a = single_stream ? 0 : seg * warp, e = single_stream ? nlit : min(a + seg, nlit);
                            const int per = (e - a + 31) / 32;
                            const int ca = min(a + lane * per, e), ce = min(ca + per, e);
                            int bits = 0;
                            for (int i = ca; i < ce; i++) bits += sm.hbits[lits[i]];
                            // encode order is last symbol first: my start = sum of bits of lanes above me
                            int incl = bits;
                            for (int o2 = 1; o2 < 32; o2 <<= 1) { int t = __shfl_down_sync(kFull, incl, o2); if (lane + o2 < 32) incl += t; }
                            uint32_t bitpos = (uint32_t) (incl - bits);
                            const int stream_bits = __shfl_sync(kFull, incl, 0);
                            uint32_t *sbuf = stage + warp * (kStreamStage / 4);
                            for (int i = ce - 1; i >= ca; i--) {
                                const uint8_t c = lits[i];
                                const int nb = sm.hbits[c];
                                or_bits(sbuf, bitpos, sm.hcode[c], nb);
                                bitpos += nb;
                            }
                            if (lane == 0) or_bits(sbuf, (uint32_t) stream_bits, 1, 1);
                        }
                        if (tid == 0) {
                            // header (:349-367)
                            const int type = 2;
                            if (lit_header == 3) { uint32_t h = type | ((single_stream ? 0 : 1) << 2) | (nlit << 4) | (total << 14); o[0] = h; o[1] = h >> 8; o[2] = h >> 16; }
                            else if (lit_header == 4) { uint32_t h = type | (2 << 2) | (nlit << 4) | ((uint32_t) total << 18); o[0] = h; o[1] = h >> 8; o[2] = h >> 16; o[3] = h >> 24; }
                            else { uint32_t h = (uint32_t) type | (3u << 2) | ((uint32_t) nlit << 4) | ((uint32_t) total << 22); o[0] = h; o[1] = h >> 8; o[2] = h >> 16; o[3] = h >> 24; o[4] = (uint8_t) ((uint32_t) total >> 10); }
                            // Huffman table description prepared in shared memory during planning
                            uint8_t *t = o + lit_header;
                            for (int i = 0; i < sm.v[V_HTABLE_BYTES]; i++) t[i] = sm.wbuf[i];
                            if (!single_stream) {
                                uint8_t *j = t + sm.v[V_HTABLE_BYTES];
                                j[0] = (uint8_t) sb0; j[1] = (uint8_t) (sb0 >> 8); j[2] = (uint8_t) sb1; j[3] = (uint8_t) (sb1 >> 8); j[4] = (uint8_t) sb2; j[5] = (uint8_t) (sb2 >> 8);
                            }
                        }
                        __syncthreads();
                        uint8_t *sdst = o + lit_header + sm.v[V_HTABLE_BYTES] + (single_stream ? 0 : 6);
                        if (warp < nstreams) {
                            int before_bytes = 0;
                            for (int k = 0; k < warp; k++) before_bytes += sm.v[V_STREAM_BYTES0 + k];
                            __syncwarp();
                            warp_copy<true>(sdst + before_bytes, reinterpret_cast<const uint8_t *>(stage + warp * (kStreamStage / 4)), sm.v[V_STREAM_BYTES0 + warp], lane);
                        }
                    }
                    else if (lit_mode == 1) {
                        if (tid == 0) {
                            if (lit_header == 1) o[0] = (uint8_t) (1 | (nlit << 3));
                            else if (lit_header == 2) { uint32_t h = 1 | (1 << 2) | (nlit << 4); o[0] = h; o[1] = h >> 8; }
                            else { uint32_t h = 1 | (3 << 2) | (nlit << 4); o[0] = h; o[1] = h >> 8; o[2] = h >> 16; }
                            o[lit_header] = lits[0];
                        }
                    }
                    else {
                        if (tid == 0) {
                            if (lit_header == 1) o[0] = (uint8_t) (0 | (nlit << 3));
                            else if (lit_header == 2) { uint32_t h = 0 | (1 << 2) | (nlit << 4); o[0] = h; o[1] = h >> 8; }
                            else { uint32_t h = 0 | (3 << 2) | (nlit << 4); o[0] = h; o[1] = h >> 8; o[2] = h >> 16; }
                        }
                        block_copy(o + lit_header, lits, nlit);
                    }
                    // ---- sequences section ----
                    uint8_t *sq = o + lit_section;
                    if (tid == 0) {
                        int p = 0;
                        if (nseq < 0x7F) sq[p++] = (uint8_t) nseq;
                        else if (nseq < 0x7F00) { sq[p++] = (uint8_t) ((nseq >> 8) | 0x80); sq[p++] = (uint8_t) nseq; }
                        else { sq[p++] = 0xFF; uint32_t v = (uint32_t) (nseq - 0x7F00); sq[p++] = (uint8_t) v; sq[p++] = (uint8_t) (v >> 8); }
                        if (nseq > 0) {
                            // encoding types (:205), then the table descriptions in the order LL, OF, ML (:96-199)
                            sq[p++] = (uint8_t) ((sm.v[V_MODE_LL] << 6) | (sm.v[V_MODE_OF] << 4) | (sm.v[V_MODE_ML] << 2));
                            for (int i = 0; i < sm.v[V_DESC_LL]; i++) sq[p++] = sm.tdesc[2][i];
                            for (int i = 0; i < sm.v[V_DESC_OF]; i++) sq[p++] = sm.tdesc[0][i];
                            for (int i = 0; i < sm.v[V_DESC_ML]; i++) sq[p++] = sm.tdesc[1][i];
                        }
                    }
                    if (nseq > 0) block_copy<true>(sq + seq_header, reinterpret_cast<const uint8_t *>(seq_stage), seq_bytes);
                    block_bytes = lit_section + seq_header + seq_bytes;
                }
            }
            if (!compressed) {
                block_copy(out + op + 3, blk, bl);
                block_bytes = bl;
            }
            if (tid == 0) {
                // 3-byte block header: last (1) | type (2) | size (21)  (ZstdFrameCompressor.writeCompressedBlock :181-204)
                const uint32_t h = (last_block ? 1u : 0u) | ((compressed ? 2u : 0u) << 1) | ((uint32_t) block_bytes << 3);
                // the decoder only sees the sequences of compressed blocks: commit the offset history for those (CompressionContext.commit)
                if (compressed) { sm.v[V_REP1] = sm.v[V_REPT1]; sm.v[V_REP2] = sm.v[V_REPT2]; sm.v[V_REP3] = sm.v[V_REPT3]; }
                out[op] = (uint8_t) h; out[op + 1] = (uint8_t) (h >> 8); out[op + 2] = (uint8_t) (h >> 16);
            }
            op += 3 + block_bytes;
            block_start += bl;
            __syncthreads();
        }
        while (!last_block);

        __syncthreads();
        if (tid == 0) {
            const uint32_t h = (uint32_t) sm.v[V_CHECKSUM];
            out[op] = (uint8_t) h; out[op + 1] = (uint8_t) (h >> 8); out[op + 2] = (uint8_t) (h >> 16); out[op + 3] = (uint8_t) (h >> 24);
            b.out_len[idx] = op + 4;
            b.status[idx] = 0;
        }
    }
}

}  // namespace

static int64_t zstd_enc_grid(int sm_count) { return (int64_t) sm_count * 4; }   // 56 KiB of shared memory and 64 registers x 256 threads per CTA

int64_t acc_zstd_enc_scratch_bytes(int sm_count, int64_t n) { return zstd_enc_grid(sm_count) * kScratchPerCta + n * 8 + 256; }

void acc_launch_zstd_compress(const AccBatch &b, int sm_count, cudaStream_t st, void *scratch, int64_t scratch_bytes)
{
    int64_t ctas = b.n;
    int64_t max_ctas = zstd_enc_grid(sm_count);
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    (void) scratch_bytes;
    // pass 1: XXH64 of every input (frame checksum, ZstdFrameCompressor.java:123-134) at HBM speed
    int64_t *hashes = reinterpret_cast<int64_t *>((uint8_t *) scratch + zstd_enc_grid(sm_count) * kScratchPerCta);
    AccBatch hb = b;
    hb.out_len = hashes;
    hb.status = nullptr;
    acc_launch_xxh64(hb, 0, sm_count, st);
    cudaFuncSetAttribute(zstd_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(EncSmem));
    zstd_compress_kernel<<<(unsigned) ctas, kThreads, sizeof(EncSmem), st>>>(b, (uint8_t *) scratch, hashes);
}
