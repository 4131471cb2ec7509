/*
 * zstd_enc_oracle.c -- CPU restatement of the reference's Zstandard frame compressor, level 3 (DFAST)
 * (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows zstd/ZstdFrameCompressor.java:52-432, zstd/CompressionParameters.java:40-145,256-324,
 * zstd/DoubleFastBlockCompressor.java:28-256, zstd/SequenceStore.java:75-159,
 * zstd/SequenceEncoder.java:66-341, zstd/FiniteStateEntropy.java:153-521, zstd/FseCompressionTable.java:41-131,
 * zstd/HuffmanCompressionTable.java:41-436, zstd/HuffmanCompressor.java:26-135, zstd/Histogram.java:28-64,
 * zstd/BitOutputStream.java:49-89, zstd/BlockCompressionState.java, zstd/RepeatedOffsets.java,
 * zstd/HuffmanCompressionContext.java.
 *
 * The public Java compressor is hard-wired to level 3 (ZstdJavaCompressor.java:51,73) and only the DFAST
 * strategy exists (CompressionParameters.java:147-183), so this file restates exactly that configuration.
 */
#include "oracle.h"
#include "zstd_oracle_common.h"
#include <stdlib.h>
#include <string.h>

#define ARG_FAIL() return ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL)
#define MIN_MATCH 3

/* ZstdJavaCompressor.maxCompressedLength :31-40 */
int64_t orc_zstd_max_compressed_length(int64_t n)
{
    int64_t r = n + (n >> 8);
    if (n < ZO_MAX_BLOCK) r += (ZO_MAX_BLOCK - n) >> 11;
    return r;
}

/* ---- CompressionParameters.compute(3, inputSize) :256-324 --------------------------------------- */
typedef struct { int window_log, chain_log, hash_log, search_log, search_length, target_length; } cparams;

static cparams compute_params(int32_t input_size)
{
    cparams p;
    /* level-3 rows of DEFAULT_COMPRESSION_PARAMETERS (:46, :72, :98, :124) */
    if (input_size <= 16 * 1024) p = (cparams) {14, 14, 14, 2, 4, 1};
    else if (input_size <= 128 * 1024) p = (cparams) {17, 15, 16, 2, 5, 1};
    else if (input_size <= 256 * 1024) p = (cparams) {18, 16, 16, 1, 4, 1};
    else p = (cparams) {20, 16, 17, 1, 5, 1};
    /* :276-283 (estimatedInputSize < 2^30 always holds for int sizes below that) */
    if ((int64_t) input_size < (1LL << 30)) {
        int input_size_log = (input_size < (1 << 6)) ? 6 : zo_highbit((uint32_t) (input_size - 1)) + 1;
        if (p.window_log > input_size_log) p.window_log = input_size_log;
    }
    if (p.hash_log > p.window_log + 1) p.hash_log = p.window_log + 1;
    int cycle_log = p.chain_log;   /* Util.cycleLog: DFAST is not a binary-tree strategy */
    if (cycle_log > p.window_log) p.chain_log -= (cycle_log - p.window_log);
    if (p.window_log < 10) p.window_log = 10;
    return p;
}

/* ---- BitOutputStream.java :49-89 ----------------------------------------------------------------- */
typedef struct { uint8_t *out; int64_t start, limit, cur; uint64_t container; int bit_count; } bitw;

static int bw_init(bitw *w, uint8_t *out, int64_t addr, int64_t size)
{
    if (size < 8) return -1;
    w->out = out; w->start = addr; w->limit = addr + size - 8; w->cur = addr; w->container = 0; w->bit_count = 0;
    return 0;
}
static inline void bw_add(bitw *w, int32_t value, int bits)
{
    w->container |= ((uint64_t) (uint32_t) value & ((1ull << bits) - 1)) << w->bit_count;   /* BIT_MASK[bits], bits <= 31 */
    w->bit_count += bits;
}
static inline void bw_add_fast(bitw *w, int32_t value, int bits) { w->container |= (uint64_t) (int64_t) value << w->bit_count; w->bit_count += bits; }
static inline void bw_flush(bitw *w)
{
    int bytes = w->bit_count >> 3;
    zo_st64(w->out + w->cur, w->container);
    w->cur += bytes;
    if (w->cur > w->limit) w->cur = w->limit;
    w->bit_count &= 7;
    w->container >>= ((bytes * 8) & 63);   /* Java masks long shift counts to 6 bits */
}
static inline int bw_close(bitw *w)
{
    bw_add_fast(w, 1, 1);
    bw_flush(w);
    if (w->cur >= w->limit) return 0;
    return (int) ((w->cur - w->start) + (w->bit_count > 0 ? 1 : 0));
}

/* ---- FseCompressionTable.java -------------------------------------------------------------------- */
typedef struct { int log2; int16_t next_state[4096]; int32_t delta_nbits[256]; int32_t delta_find[256]; } fse_ctable;

static void fse_ct_rle(fse_ctable *t, int symbol)     /* :41-50 */
{
    t->log2 = 0; t->next_state[0] = 0; t->next_state[1] = 0; t->delta_find[symbol] = 0; t->delta_nbits[symbol] = 0;
}

static void fse_ct_init(fse_ctable *t, const int16_t *norm, int max_symbol, int table_log)   /* :52-111 */
{
    int size = 1 << table_log;
    uint8_t table[4096];
    int cumulative[258];
    int high = size - 1;
    t->log2 = table_log;
    cumulative[0] = 0;
    for (int i = 1; i <= max_symbol + 1; i++) {
        if (norm[i - 1] == -1) { cumulative[i] = cumulative[i - 1] + 1; table[high--] = (uint8_t) (i - 1); }
        else cumulative[i] = cumulative[i - 1] + norm[i - 1];
    }
    cumulative[max_symbol + 1] = size + 1;
    zo_spread_symbols(norm, max_symbol, size, high, table);
    for (int i = 0; i < size; i++) { int s = table[i]; t->next_state[cumulative[s]++] = (int16_t) (size + i); }
    int total = 0;
    for (int s = 0; s <= max_symbol; s++) {
        int n = norm[s];
        if (n == 0) t->delta_nbits[s] = ((table_log + 1) << 16) - size;
        else if (n == -1 || n == 1) { t->delta_nbits[s] = (table_log << 16) - size; t->delta_find[s] = total - 1; total++; }
        else {
            int max_bits_out = table_log - zo_highbit((uint32_t) (n - 1));
            int min_state_plus = n << max_bits_out;
            t->delta_nbits[s] = (max_bits_out << 16) - min_state_plus;
            t->delta_find[s] = total - n;
            total += n;
        }
    }
}
static inline int fse_ct_begin(const fse_ctable *t, int symbol)   /* :113-118 */
{
    int output_bits = (int) ((uint32_t) (t->delta_nbits[symbol] + (1 << 15)) >> 16);
    int base = (int) ((uint32_t) ((output_bits << 16) - t->delta_nbits[symbol]) >> output_bits);
    return t->next_state[base + t->delta_find[symbol]];
}
static inline int fse_ct_encode(const fse_ctable *t, bitw *w, int state, int symbol)   /* :120-125 */
{
    int output_bits = (int) ((uint32_t) (state + t->delta_nbits[symbol]) >> 16);
    bw_add(w, state, output_bits);
    return t->next_state[(int) ((uint32_t) state >> output_bits) + t->delta_find[symbol]];
}
static inline void fse_ct_finish(const fse_ctable *t, bitw *w, int state) { bw_add(w, state, t->log2); bw_flush(w); }   /* :127-131 */

/* ---- FiniteStateEntropy.java (encode half) ------------------------------------------------------- */
static int min_table_log(int input_size, int max_symbol)   /* Util.minTableLog :120-131 */
{
    int a = zo_highbit((uint32_t) (input_size - 1)) + 1, b = zo_highbit((uint32_t) max_symbol) + 2;
    return a < b ? a : b;
}
static int fse_optimal_table_log(int max_table_log, int input_size, int max_symbol)   /* :238-255 */
{
    int r = max_table_log, v = zo_highbit((uint32_t) (input_size - 1)) - 2;
    if (v < r) r = v;
    v = min_table_log(input_size, max_symbol);
    if (v > r) r = v;
    if (r < 5) r = 5;
    if (r > 12) r = 12;
    return r;
}

static const int REST_TO_BEAT[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};

static void fse_normalize2(int16_t *norm, int table_log, const int32_t *counts, int total, int max_symbol)   /* :315-405 */
{
    const int16_t UNASSIGNED = -2;
    int distributed = 0;
    int low_threshold = (int) ((uint32_t) total >> table_log);
    int low_one = (int) ((uint32_t) (total * 3) >> (table_log + 1));
    for (int i = 0; i <= max_symbol; i++) {
        if (counts[i] == 0) norm[i] = 0;
        else if (counts[i] <= low_threshold) { norm[i] = -1; distributed++; total -= counts[i]; }
        else if (counts[i] <= low_one) { norm[i] = 1; distributed++; total -= counts[i]; }
        else norm[i] = UNASSIGNED;
    }
    int factor = 1 << table_log;
    int to_distribute = factor - distributed;
    if ((total / to_distribute) > low_one) {
        low_one = (total * 3) / (to_distribute * 2);
        for (int i = 0; i <= max_symbol; i++) {
            if (norm[i] == UNASSIGNED && counts[i] <= low_one) { norm[i] = 1; distributed++; total -= counts[i]; }
        }
        to_distribute = factor - distributed;
    }
    if (distributed == max_symbol + 1) {
        int max_value = 0, max_count = 0;
        for (int i = 0; i <= max_symbol; i++) if (counts[i] > max_count) { max_value = i; max_count = counts[i]; }
        norm[max_value] = (int16_t) (norm[max_value] + (int16_t) to_distribute);
        return;
    }
    if (total == 0) {
        for (int i = 0; to_distribute > 0; i = (i + 1) % (max_symbol + 1)) {
            if (norm[i] > 0) { to_distribute--; norm[i]++; }
        }
        return;
    }
    int64_t v_step_log = 62 - table_log;
    int64_t mid = (1LL << (v_step_log - 1)) - 1;
    int64_t r_step = (((1LL << v_step_log) * to_distribute) + mid) / total;
    int64_t tmp_total = mid;
    for (int i = 0; i <= max_symbol; i++) {
        if (norm[i] == UNASSIGNED) {
            int64_t end = tmp_total + ((int64_t) counts[i] * r_step);
            int s_start = (int) ((uint64_t) tmp_total >> v_step_log);
            int s_end = (int) ((uint64_t) end >> v_step_log);
            norm[i] = (int16_t) (s_end - s_start);
            tmp_total = end;
        }
    }
}

static void fse_normalize(int16_t *norm, int table_log, const int32_t *counts, int total, int max_symbol)   /* :257-313 */
{
    int64_t scale = 62 - table_log;
    int64_t step = (1LL << 62) / total;
    int64_t vstep = 1LL << (scale - 20);
    int still = 1 << table_log;
    int largest = 0;
    int16_t largest_p = 0;
    int low_threshold = (int) ((uint32_t) total >> table_log);
    for (int s = 0; s <= max_symbol; s++) {
        if (counts[s] == 0) { norm[s] = 0; continue; }
        if (counts[s] <= low_threshold) { norm[s] = -1; still--; }
        else {
            int16_t p = (int16_t) ((uint64_t) ((int64_t) counts[s] * step) >> scale);
            if (p < 8) {
                int64_t rest_to_beat = vstep * REST_TO_BEAT[p];
                int64_t delta = (int64_t) counts[s] * step - (((int64_t) p) << scale);
                if (delta > rest_to_beat) p++;
            }
            if (p > largest_p) { largest_p = p; largest = s; }
            norm[s] = p;
            still -= p;
        }
    }
    if (-still >= (int) ((uint32_t) (int32_t) norm[largest] >> 1)) fse_normalize2(norm, table_log, counts, total, max_symbol);
    else norm[largest] = (int16_t) (norm[largest] + (int16_t) still);
}

/* writeNormalizedCounts :407-521; returns size or -1 */
static int fse_write_ncount(uint8_t *out, int64_t addr, int64_t out_size, const int16_t *norm, int max_symbol, int table_log)
{
    int64_t output = addr, limit = addr + out_size;
    int table_size = 1 << table_log;
    int bit_count = 0;
    int32_t bit_stream = table_log - 5;
    bit_count += 4;
    int remaining = table_size + 1, threshold = table_size, table_bits = table_log + 1;
    int symbol = 0, previous0 = 0;
    while (remaining > 1) {
        if (previous0) {
            int start = symbol;
            while (norm[symbol] == 0) symbol++;
            while (symbol >= start + 24) {
                start += 24;
                bit_stream |= (int32_t) (0xFFFFu << bit_count);
                if (output + 2 > limit) return -1;
                zo_st16(out + output, (uint32_t) bit_stream);
                output += 2;
                bit_stream = (int32_t) ((uint32_t) bit_stream >> 16);
            }
            while (symbol >= start + 3) { start += 3; bit_stream |= 3 << bit_count; bit_count += 2; }
            bit_stream |= (symbol - start) << bit_count;
            bit_count += 2;
            if (bit_count > 16) {
                if (output + 2 > limit) return -1;
                zo_st16(out + output, (uint32_t) bit_stream);
                output += 2;
                bit_stream = (int32_t) ((uint32_t) bit_stream >> 16);
                bit_count -= 16;
            }
        }
        int count = norm[symbol++];
        int max = (2 * threshold - 1) - remaining;
        remaining -= count < 0 ? -count : count;
        count++;
        if (count >= threshold) count += max;
        bit_stream |= (int32_t) ((uint32_t) count << bit_count);
        bit_count += table_bits;
        bit_count -= (count < max ? 1 : 0);
        previous0 = (count == 1);
        while (remaining < threshold) { table_bits--; threshold >>= 1; }
        if (bit_count > 16) {
            if (output + 2 > limit) return -1;
            zo_st16(out + output, (uint32_t) bit_stream);
            output += 2;
            bit_stream = (int32_t) ((uint32_t) bit_stream >> 16);
            bit_count -= 16;
        }
    }
    if (output + 2 > limit) return -1;
    zo_st16(out + output, (uint32_t) bit_stream);
    output += (bit_count + 7) / 8;
    return (int) (output - addr);
}

/* FiniteStateEntropy.compress :153-236 (used for Huffman weights) */
static int fse_compress(uint8_t *out, int64_t addr, int64_t out_size, const uint8_t *in, int in_size, const fse_ctable *t)
{
    if (out_size < 8) return -1;
    int input = in_size;
    if (in_size <= 2) return 0;
    bitw w;
    bw_init(&w, out, addr, out_size);
    int state1, state2;
    if (in_size & 1) {
        state1 = fse_ct_begin(t, in[--input]);
        state2 = fse_ct_begin(t, in[--input]);
        state1 = fse_ct_encode(t, &w, state1, in[--input]);
        bw_flush(&w);
    }
    else {
        state2 = fse_ct_begin(t, in[--input]);
        state1 = fse_ct_begin(t, in[--input]);
    }
    in_size -= 2;
    if (in_size & 2) {   /* 64 > 12*4+7 */
        state2 = fse_ct_encode(t, &w, state2, in[--input]);
        state1 = fse_ct_encode(t, &w, state1, in[--input]);
        bw_flush(&w);
    }
    while (input > 0) {
        state2 = fse_ct_encode(t, &w, state2, in[--input]);
        state1 = fse_ct_encode(t, &w, state1, in[--input]);
        state2 = fse_ct_encode(t, &w, state2, in[--input]);
        state1 = fse_ct_encode(t, &w, state1, in[--input]);
        bw_flush(&w);
    }
    fse_ct_finish(t, &w, state2);
    fse_ct_finish(t, &w, state1);
    return bw_close(&w);
}

/* ---- Histogram.java ------------------------------------------------------------------------------ */
static void histogram(const uint8_t *in, int n, int32_t *counts, int nsym)
{
    memset(counts, 0, sizeof(int32_t) * (size_t) nsym);
    for (int i = 0; i < n; i++) counts[in[i]]++;
}
static int find_max_symbol(const int32_t *counts, int max_symbol) { while (counts[max_symbol] == 0) max_symbol--; return max_symbol; }
static int find_largest(const int32_t *counts, int max_symbol) { int m = 0; for (int i = 0; i <= max_symbol; i++) if (counts[i] > m) m = counts[i]; return m; }

/* ---- HuffmanCompressionTable.java ---------------------------------------------------------------- */
typedef struct { int16_t values[256]; uint8_t nbits[256]; int max_symbol, max_nbits; } huf_ctable;
typedef struct { int32_t count[512]; int16_t parents[512]; int32_t symbols[512]; uint8_t nbits[512]; } node_table;

static int huf_optimal_bits(int max_bits, int input_size, int max_symbol)   /* :41-58 */
{
    int r = max_bits, v = zo_highbit((uint32_t) (input_size - 1)) - 1;
    if (v < r) r = v;
    v = min_table_log(input_size, max_symbol);
    if (v > r) r = v;
    if (r < 5) r = 5;
    if (r > 12) r = 12;
    return r;
}

static int huf_build_tree(const int32_t *counts, int max_symbol, node_table *nt)   /* :105-190 */
{
    int current = 0;
    for (int symbol = 0; symbol <= max_symbol; symbol++) {
        int count = counts[symbol];
        int position = current;
        while (position > 1 && count > nt->count[position - 1]) {
            nt->count[position] = nt->count[position - 1]; nt->parents[position] = nt->parents[position - 1];
            nt->symbols[position] = nt->symbols[position - 1]; nt->nbits[position] = nt->nbits[position - 1];
            position--;
        }
        nt->count[position] = count;
        nt->symbols[position] = symbol;
        current++;
    }
    int last_non_zero = max_symbol;
    while (nt->count[last_non_zero] == 0) last_non_zero--;
    const int non_leaf_start = 256;
    current = non_leaf_start;
    int current_leaf = last_non_zero;
    int current_non_leaf = current;
    nt->count[current] = nt->count[current_leaf] + nt->count[current_leaf - 1];
    nt->parents[current_leaf] = (int16_t) current;
    nt->parents[current_leaf - 1] = (int16_t) current;
    current++;
    current_leaf -= 2;
    int root = 256 + last_non_zero - 1;
    for (int n = current; n <= root; n++) nt->count[n] = 1 << 30;
    while (current <= root) {
        int child1, child2;
        if (current_leaf >= 0 && nt->count[current_leaf] < nt->count[current_non_leaf]) child1 = current_leaf--; else child1 = current_non_leaf++;
        if (current_leaf >= 0 && nt->count[current_leaf] < nt->count[current_non_leaf]) child2 = current_leaf--; else child2 = current_non_leaf++;
        nt->count[current] = nt->count[child1] + nt->count[child2];
        nt->parents[child1] = (int16_t) current;
        nt->parents[child2] = (int16_t) current;
        current++;
    }
    nt->nbits[root] = 0;
    for (int n = root - 1; n >= non_leaf_start; n--) nt->nbits[n] = (uint8_t) (nt->nbits[nt->parents[n]] + 1);
    for (int n = 0; n <= last_non_zero; n++) nt->nbits[n] = (uint8_t) (nt->nbits[nt->parents[n]] + 1);
    return last_non_zero;
}

static int huf_set_max_height(node_table *nt, int last_non_zero, int max_bits)   /* :294-390 */
{
    int largest_bits = nt->nbits[last_non_zero];
    if (largest_bits <= max_bits) return largest_bits;
    int total_cost = 0;
    int base_cost = 1 << (largest_bits - max_bits);
    int n = last_non_zero;
    while (nt->nbits[n] > max_bits) {
        total_cost += base_cost - (1 << (largest_bits - nt->nbits[n]));
        nt->nbits[n] = (uint8_t) max_bits;
        n--;
    }
    while (nt->nbits[n] == max_bits) n--;
    total_cost = (int) ((uint32_t) total_cost >> (largest_bits - max_bits));
    const int no_symbol = (int) 0xF0F0F0F0;
    int rank_last[14];
    for (int i = 0; i < 14; i++) rank_last[i] = no_symbol;
    int current_bits = max_bits;
    for (int pos = n; pos >= 0; pos--) {
        if (nt->nbits[pos] >= current_bits) continue;
        current_bits = nt->nbits[pos];
        rank_last[max_bits - current_bits] = pos;
    }
    while (total_cost > 0) {
        int dec = zo_highbit((uint32_t) total_cost) + 1;
        for (; dec > 1; dec--) {
            int high_pos = rank_last[dec], low_pos = rank_last[dec - 1];
            if (high_pos == no_symbol) continue;
            if (low_pos == no_symbol) break;
            int high_total = nt->count[high_pos], low_total = 2 * nt->count[low_pos];
            if (high_total <= low_total) break;
        }
        while (dec <= 12 && rank_last[dec] == no_symbol) dec++;
        total_cost -= 1 << (dec - 1);
        if (rank_last[dec - 1] == no_symbol) rank_last[dec - 1] = rank_last[dec];
        nt->nbits[rank_last[dec]]++;
        if (rank_last[dec] == 0) rank_last[dec] = no_symbol;
        else {
            rank_last[dec]--;
            if (nt->nbits[rank_last[dec]] != max_bits - dec) rank_last[dec] = no_symbol;
        }
    }
    while (total_cost < 0) {
        if (rank_last[1] == no_symbol) {
            while (nt->nbits[n] == max_bits) n--;
            nt->nbits[n + 1]--;
            rank_last[1] = n + 1;
            total_cost++;
            continue;
        }
        nt->nbits[rank_last[1] + 1]--;
        rank_last[1]++;
        total_cost++;
    }
    return max_bits;
}

static void huf_ct_init(huf_ctable *t, const int32_t *counts, int max_symbol, int max_bits)   /* :60-103 */
{
    node_table nt;
    memset(&nt, 0, sizeof(nt));
    int last_non_zero = huf_build_tree(counts, max_symbol, &nt);
    max_bits = huf_set_max_height(&nt, last_non_zero, max_bits);
    for (int node = 0; node <= max_symbol; node++) t->nbits[nt.symbols[node]] = nt.nbits[node];
    int16_t entries[13], values[13];
    memset(entries, 0, sizeof(entries));
    memset(values, 0, sizeof(values));
    for (int n = 0; n <= last_non_zero; n++) entries[nt.nbits[n]]++;
    int16_t starting = 0;
    for (int rank = max_bits; rank > 0; rank--) {
        values[rank] = starting;
        starting = (int16_t) (starting + entries[rank]);
        starting = (int16_t) ((uint32_t) (int32_t) starting >> 1);
    }
    for (int n = 0; n <= max_symbol; n++) t->values[n] = values[t->nbits[n]]++;
    t->max_symbol = max_symbol;
    t->max_nbits = max_bits;
}

/* compressWeights :395-436; returns size, 0 = not compressible, 1 = single symbol */
static int huf_compress_weights(uint8_t *out, int64_t addr, int64_t out_size, const uint8_t *weights, int n)
{
    if (n <= 1) return 0;
    int32_t counts[13];
    histogram(weights, n, counts, 13);
    int max_symbol = find_max_symbol(counts, 12);
    int max_count = find_largest(counts, max_symbol);
    if (max_count == n) return 1;
    if (max_count == 1) return 0;
    int16_t norm[13];
    int table_log = fse_optimal_table_log(6, n, max_symbol);
    fse_normalize(norm, table_log, counts, n, max_symbol);
    int64_t output = addr, limit = addr + out_size;
    int hs = fse_write_ncount(out, output, out_size, norm, max_symbol, table_log);
    if (hs < 0) return -1;
    output += hs;
    fse_ctable ct;
    fse_ct_init(&ct, norm, max_symbol, table_log);
    int cs = fse_compress(out, output, limit - output, weights, n, &ct);
    if (cs < 0) return -1;
    if (cs == 0) return 0;
    output += cs;
    return (int) (output - addr);
}

static int huf_ct_write(const huf_ctable *t, uint8_t *out, int64_t addr, int64_t out_size)   /* :202-263 */
{
    uint8_t weights[256];
    int64_t output = addr;
    int max_symbol = t->max_symbol;
    for (int s = 0; s < max_symbol; s++) weights[s] = t->nbits[s] == 0 ? 0 : (uint8_t) (t->max_nbits + 1 - t->nbits[s]);
    int size = huf_compress_weights(out, output + 1, out_size - 1, weights, max_symbol);
    if (size < 0) return -1;
    if (size != 0 && size != 1 && size < max_symbol / 2) { out[output] = (uint8_t) size; return size + 1; }
    int entry_count = max_symbol;
    size = (entry_count + 1) / 2;
    if (size + 1 > out_size) return -1;
    out[output++] = (uint8_t) (127 + entry_count);
    weights[max_symbol] = 0;
    for (int i = 0; i < entry_count; i += 2) out[output++] = (uint8_t) ((weights[i] << 4) + weights[i + 1]);
    return (int) (output - addr);
}

static int huf_ct_estimate(const huf_ctable *t, const int32_t *counts, int max_symbol)   /* :283-291 */
{
    int bits = 0, m = max_symbol < t->max_symbol ? max_symbol : t->max_symbol;
    for (int s = 0; s <= m; s++) bits += t->nbits[s] * counts[s];
    return (int) ((uint32_t) bits >> 3);
}
static int huf_ct_valid(const huf_ctable *t, const int32_t *counts, int max_symbol)   /* :268-281 */
{
    if (max_symbol > t->max_symbol) return 0;
    for (int s = 0; s <= max_symbol; ++s) if (counts[s] != 0 && t->nbits[s] == 0) return 0;
    return 1;
}

/* HuffmanCompressor.compressSingleStream :82-135 */
static int huf_compress_1(uint8_t *out, int64_t addr, int64_t out_size, const uint8_t *in, int in_size, const huf_ctable *t)
{
    if (out_size < 8) return 0;
    bitw w;
    bw_init(&w, out, addr, out_size);
    int n = in_size & ~3;
#define HENC(sym) bw_add_fast(&w, t->values[sym], t->nbits[sym])
    switch (in_size & 3) {
        case 3: HENC(in[n + 2]); /* fall through */
        case 2: HENC(in[n + 1]); /* fall through */
        case 1: HENC(in[n + 0]); bw_flush(&w); /* fall through */
        default: break;
    }
    for (; n > 0; n -= 4) {
        HENC(in[n - 1]); HENC(in[n - 2]); HENC(in[n - 3]); HENC(in[n - 4]);
        bw_flush(&w);
    }
#undef HENC
    return bw_close(&w);
}

/* HuffmanCompressor.compress4streams :26-80 */
static int huf_compress_4(uint8_t *out, int64_t addr, int64_t out_size, const uint8_t *in, int in_size, const huf_ctable *t)
{
    int64_t output = addr, limit = addr + out_size;
    int segment = (in_size + 3) / 4;
    if (out_size < 6 + 1 + 1 + 1 + 8) return 0;
    if (in_size <= 6 + 1 + 1 + 1) return 0;
    output += 6;
    const uint8_t *p = in;
    for (int k = 0; k < 3; k++) {
        int cs = huf_compress_1(out, output, limit - output, p, segment, t);
        if (cs == 0) return 0;
        zo_st16(out + addr + 2 * k, (uint32_t) cs);
        output += cs;
        p += segment;
    }
    int cs = huf_compress_1(out, output, limit - output, p, (int) (in + in_size - p), t);
    if (cs == 0) return 0;
    output += cs;
    return (int) (output - addr);
}

/* ---- compression context (CompressionContext.java, HuffmanCompressionContext.java, RepeatedOffsets.java) */
typedef struct {
    cparams p;
    int32_t *hash_table, *chain_table;
    int32_t window_base_offset;
    int32_t rep0, rep1, tmp0, tmp1;
    /* SequenceStore */
    uint8_t *literals; int32_t literals_len;
    int32_t *offsets, *lit_lengths, *match_lengths;
    uint8_t *ll_codes, *ml_codes, *of_codes;
    int32_t seq_count;
    int long_field;   /* 0 none, 1 literal, 2 match */
    int32_t long_pos;
    /* Huffman tables */
    huf_ctable huf_a, huf_b;
    huf_ctable *prev_table, *temp_table, *prev_cand, *temp_cand;
    fse_ctable ll_ct, of_ct, ml_ct;
    fse_ctable def_ll, def_of, def_ml;
} zcctx;

static void store_sequence(zcctx *c, const uint8_t *in, int64_t lit_addr, int32_t lit_len, int32_t offset_code, int32_t ml_base)   /* SequenceStore :81-112 */
{
    memcpy(c->literals + c->literals_len, in + lit_addr, (size_t) lit_len);   /* the Java wild-copies; kept bytes identical */
    c->literals_len += lit_len;
    if (lit_len > 65535) { c->long_field = 1; c->long_pos = c->seq_count; }
    c->lit_lengths[c->seq_count] = lit_len;
    c->offsets[c->seq_count] = offset_code + 1;
    if (ml_base > 65535) { c->long_field = 2; c->long_pos = c->seq_count; }
    c->match_lengths[c->seq_count] = ml_base;
    c->seq_count++;
}

static const uint8_t LL_CODE[64] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21,
                                    22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24};
static const uint8_t ML_CODE[128] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
                                     32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 36, 37, 37, 37, 37, 38, 38, 38, 38, 38, 38, 38, 38, 39, 39, 39, 39, 39, 39, 39, 39,
                                     40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41,
                                     42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42};

static void generate_codes(zcctx *c)   /* SequenceStore :121-159 */
{
    for (int i = 0; i < c->seq_count; ++i) {
        int32_t ll = c->lit_lengths[i], ml = c->match_lengths[i];
        c->ll_codes[i] = (uint8_t) (ll >= 64 ? zo_highbit((uint32_t) ll) + 19 : LL_CODE[ll]);
        c->of_codes[i] = (uint8_t) zo_highbit((uint32_t) c->offsets[i]);
        c->ml_codes[i] = (uint8_t) (ml >= 128 ? zo_highbit((uint32_t) ml) + 36 : ML_CODE[ml]);
    }
    if (c->long_field == 1) c->ll_codes[c->long_pos] = 35;
    if (c->long_field == 2) c->ml_codes[c->long_pos] = 52;
}

/* ---- DoubleFastBlockCompressor.java :28-256 ------------------------------------------------------- */
static inline int hash4(uint32_t v, int bits) { return (int) ((v * 0x9E3779B1u) >> (32 - bits)); }
static inline int hash5(uint64_t v, int bits) { return (int) (((v << 24) * 0xCF1BBCDCBBULL) >> (64 - bits)); }
static inline int hash6(uint64_t v, int bits) { return (int) (((v << 16) * 0xCF1BBCDCBF9BULL) >> (64 - bits)); }
static inline int hash7(uint64_t v, int bits) { return (int) (((v << 8) * 0xCF1BBCDCBFA563ULL) >> (64 - bits)); }
static inline int hash8(uint64_t v, int bits) { return (int) ((v * 0xCF1BBCDCB7A56463ULL) >> (64 - bits)); }
static inline int hash_n(const uint8_t *in, int64_t addr, int bits, int len)
{
    switch (len) {
        case 8: return hash8(zo_ld64(in + addr), bits);
        case 7: return hash7(zo_ld64(in + addr), bits);
        case 6: return hash6(zo_ld64(in + addr), bits);
        case 5: return hash5(zo_ld64(in + addr), bits);
        default: return hash4(zo_ld32(in + addr), bits);
    }
}
static int count_match(const uint8_t *in, int64_t input, int64_t limit, int64_t match)   /* :187-214 */
{
    int remaining = (int) (limit - input), count = 0;
    while (count < remaining - 7) {
        uint64_t diff = zo_ld64(in + match) ^ zo_ld64(in + input);
        if (diff != 0) return count + (__builtin_ctzll(diff) >> 3);
        count += 8; input += 8; match += 8;
    }
    while (count < remaining && in[match] == in[input]) { count++; input++; match++; }
    return count;
}

/* `in` is the frame base (baseAddress = 0); block = [block_start, block_start + size) */
static int32_t dfast_compress_block(zcctx *c, const uint8_t *in, int64_t block_start, int32_t size)
{
    const int msl = c->p.search_length > 4 ? c->p.search_length : 4;
    const int64_t window_base = c->window_base_offset;
    int32_t *long_t = c->hash_table, *short_t = c->chain_table;
    const int long_bits = c->p.hash_log, short_bits = c->p.chain_log;
    const int64_t input_end = block_start + size;
    const int64_t input_limit = input_end - 8;
    int64_t input = block_start, anchor = block_start;
    int32_t offset1 = c->rep0, offset2 = c->rep1, saved = 0;
    if (input - window_base == 0) input++;
    int32_t max_rep = (int32_t) (input - window_base);
    if (offset2 > max_rep) { saved = offset2; offset2 = 0; }
    if (offset1 > max_rep) { saved = offset1; offset1 = 0; }

    while (input < input_limit) {
        int sh = hash_n(in, input, short_bits, msl);
        int64_t short_match = short_t[sh];
        int lh = hash8(zo_ld64(in + input), long_bits);
        int64_t long_match = long_t[lh];
        int32_t current = (int32_t) input;
        long_t[lh] = current;
        short_t[sh] = current;
        int32_t match_length, offset;
        if (offset1 > 0 && zo_ld32(in + input + 1 - offset1) == zo_ld32(in + input + 1)) {
            match_length = count_match(in, input + 1 + 4, input_end, input + 1 + 4 - offset1) + 4;
            input++;
            store_sequence(c, in, anchor, (int32_t) (input - anchor), 0, match_length - MIN_MATCH);
        }
        else {
            if (long_match > window_base && zo_ld64(in + long_match) == zo_ld64(in + input)) {
                match_length = count_match(in, input + 8, input_end, long_match + 8) + 8;
                offset = (int32_t) (input - long_match);
                while (input > anchor && long_match > window_base && in[input - 1] == in[long_match - 1]) { input--; long_match--; match_length++; }
            }
            else if (short_match > window_base && zo_ld32(in + short_match) == zo_ld32(in + input)) {
                int nh = hash8(zo_ld64(in + input + 1), long_bits);
                int64_t next_match = long_t[nh];
                long_t[nh] = current + 1;
                if (next_match > window_base && zo_ld64(in + next_match) == zo_ld64(in + input + 1)) {
                    match_length = count_match(in, input + 1 + 8, input_end, next_match + 8) + 8;
                    input++;
                    offset = (int32_t) (input - next_match);
                    while (input > anchor && next_match > window_base && in[input - 1] == in[next_match - 1]) { input--; next_match--; match_length++; }
                }
                else {
                    match_length = count_match(in, input + 4, input_end, short_match + 4) + 4;
                    offset = (int32_t) (input - short_match);
                    while (input > anchor && short_match > window_base && in[input - 1] == in[short_match - 1]) { input--; short_match--; match_length++; }
                }
            }
            else {
                input += ((input - anchor) >> 8) + 1;
                continue;
            }
            offset2 = offset1;
            offset1 = offset;
            store_sequence(c, in, anchor, (int32_t) (input - anchor), offset + 2, match_length - MIN_MATCH);
        }
        input += match_length;
        anchor = input;
        if (input <= input_limit) {
            long_t[hash8(zo_ld64(in + current + 2), long_bits)] = current + 2;
            short_t[hash_n(in, current + 2, short_bits, msl)] = current + 2;
            long_t[hash8(zo_ld64(in + input - 2), long_bits)] = (int32_t) (input - 2);
            short_t[hash_n(in, input - 2, short_bits, msl)] = (int32_t) (input - 2);
            while (input <= input_limit && offset2 > 0 && zo_ld32(in + input) == zo_ld32(in + input - offset2)) {
                int32_t rep_len = count_match(in, input + 4, input_end, input + 4 - offset2) + 4;
                int32_t t = offset2; offset2 = offset1; offset1 = t;
                short_t[hash_n(in, input, short_bits, msl)] = (int32_t) input;
                long_t[hash8(zo_ld64(in + input), long_bits)] = (int32_t) input;
                store_sequence(c, in, anchor, 0, 0, rep_len - MIN_MATCH);
                input += rep_len;
                anchor = input;
            }
        }
    }
    c->tmp0 = offset1 != 0 ? offset1 : saved;
    c->tmp1 = offset2 != 0 ? offset2 : saved;
    return (int32_t) (input_end - anchor);
}

/* ---- literals: ZstdFrameCompressor.encodeLiterals :262-378, rawLiterals :407-432, rleLiterals :380-398 */
static int raw_literals(uint8_t *out, int64_t addr, int64_t out_size, const uint8_t *lit, int n)
{
    int hs = 1 + (n >= 32) + (n >= 4096);
    if (n + hs > out_size) return -1;
    if (hs == 1) out[addr] = (uint8_t) (0 | (n << 3));
    else if (hs == 2) zo_st16(out + addr, (uint32_t) (0 | (1 << 2) | (n << 4)));
    else zo_st24(out + addr, (uint32_t) (0 | (3 << 2) | (n << 4)));
    memcpy(out + addr + hs, lit, (size_t) n);
    return hs + n;
}
static int rle_literals(uint8_t *out, int64_t addr, const uint8_t *lit, int n)
{
    int hs = 1 + (n > 31) + (n > 4095);
    if (hs == 1) out[addr] = (uint8_t) (1 | (n << 3));
    else if (hs == 2) zo_st16(out + addr, (uint32_t) (1 | (1 << 2) | (n << 4)));
    else zo_st32(out + addr, (uint32_t) (1 | (3 << 2) | (n << 4)));   /* the Java writes 4 bytes here; the 4th is overwritten below or by later output */
    out[addr + hs] = lit[0];
    return hs + 1;
}
static int min_gain(int n) { return (int) ((uint32_t) n >> 6) + 2; }   /* calculateMinimumGain :400-405 (not BTULTRA) */

static int encode_literals(zcctx *c, uint8_t *out, int64_t addr, int64_t out_size, const uint8_t *lit, int n)
{
    if (n <= 63) return raw_literals(out, addr, out_size, lit, n);    /* MINIMUM_LITERALS_SIZE */
    int hs = 3 + (n >= 1024) + (n >= 16384);
    if (hs + 1 > out_size) return -1;
    int32_t counts[256];
    histogram(lit, n, counts, 256);
    int max_symbol = find_max_symbol(counts, 255);
    int largest = find_largest(counts, max_symbol);
    if (largest == n) return rle_literals(out, addr, lit, n);
    if (largest <= (int) ((uint32_t) n >> 7) + 4) return raw_literals(out, addr, out_size, lit, n);

    huf_ctable *previous = c->prev_table, *table;
    int ser_size, reuse;
    int can_reuse = huf_ct_valid(previous, counts, max_symbol);
    int prefer_reuse = n <= 1024;   /* strategy DFAST < LAZY */
    if (prefer_reuse && can_reuse) { table = previous; reuse = 1; ser_size = 0; }
    else {
        huf_ctable *nt = c->temp_table;                        /* borrowTemporaryTable */
        c->prev_cand = c->temp_table; c->temp_cand = c->prev_table;
        memset(nt->nbits, 0, sizeof(nt->nbits));
        huf_ct_init(nt, counts, max_symbol, huf_optimal_bits(11, n, max_symbol));
        ser_size = huf_ct_write(nt, out, addr + hs, out_size - hs);
        if (ser_size < 0) return -1;
        if (can_reuse && huf_ct_estimate(previous, counts, max_symbol) <= ser_size + huf_ct_estimate(nt, counts, max_symbol)) {
            table = previous; reuse = 1; ser_size = 0;
            c->prev_cand = c->prev_table; c->temp_cand = c->temp_table;   /* discardTemporaryTable */
        }
        else { table = nt; reuse = 0; }
    }
    int single = n < 256;
    int cs = single ? huf_compress_1(out, addr + hs + ser_size, out_size - hs - ser_size, lit, n, table)
                    : huf_compress_4(out, addr + hs + ser_size, out_size - hs - ser_size, lit, n, table);
    int total = ser_size + cs;
    if (cs == 0 || total >= n - min_gain(n)) {
        c->prev_cand = c->prev_table; c->temp_cand = c->temp_table;
        return raw_literals(out, addr, out_size, lit, n);
    }
    int type = reuse ? 3 : 2;
    if (hs == 3) zo_st24(out + addr, (uint32_t) (type | ((single ? 0 : 1) << 2) | (n << 4) | (total << 14)));
    else if (hs == 4) zo_st32(out + addr, (uint32_t) (type | (2 << 2) | (n << 4) | (total << 18)));
    else { zo_st32(out + addr, (uint32_t) type | (3u << 2) | ((uint32_t) n << 4) | ((uint32_t) total << 22)); out[addr + 4] = (uint8_t) ((uint32_t) total >> 10); }
    return hs + total;
}

/* ---- sequences: SequenceEncoder.java ---------------------------------------------------------------- */
static int select_encoding(int largest, int seq_count, int def_log, int default_allowed)   /* :299-341, strategy ordinal 1 (DFAST) */
{
    if (largest == seq_count) {
        if (default_allowed && seq_count <= 2) return 0;
        return 1;
    }
    if (default_allowed) {
        int64_t min_seq = ((1LL << def_log) * 9) >> 3;
        if (seq_count < min_seq || largest < (seq_count >> (def_log - 1))) return 0;
    }
    return 2;
}

static int build_ctable(fse_ctable *t, uint8_t *out, int64_t output, int64_t limit, int seq_count, int max_table_log, const uint8_t *codes,
                        int32_t *counts, int max_symbol)   /* buildCompressionTable :211-226 */
{
    int16_t norm[64];
    int table_log = fse_optimal_table_log(max_table_log, seq_count, max_symbol);
    if (counts[codes[seq_count - 1]] > 1) { counts[codes[seq_count - 1]]--; seq_count--; }
    fse_normalize(norm, table_log, counts, seq_count, max_symbol);
    fse_ct_init(t, norm, max_symbol, table_log);
    return fse_write_ncount(out, output, limit - output, norm, max_symbol, table_log);
}

static int encode_sequences(zcctx *c, uint8_t *out, int64_t output, int64_t limit, const fse_ctable *mlt, const fse_ctable *oft, const fse_ctable *llt)   /* :228-297 */
{
    bitw w;
    if (bw_init(&w, out, output, limit - output) != 0) return -1;
    int n = c->seq_count;
    const uint8_t *mlc = c->ml_codes, *ofc = c->of_codes, *llc = c->ll_codes;
    int ml_state = fse_ct_begin(mlt, mlc[n - 1]);
    int of_state = fse_ct_begin(oft, ofc[n - 1]);
    int ll_state = fse_ct_begin(llt, llc[n - 1]);
    bw_add(&w, c->lit_lengths[n - 1], ZO_LL_BITS[llc[n - 1]]);
    bw_add(&w, c->match_lengths[n - 1], ZO_ML_BITS[mlc[n - 1]]);
    bw_add(&w, c->offsets[n - 1], ofc[n - 1]);
    bw_flush(&w);
    for (int i = n - 2; i >= 0; i--) {
        int ll_code = llc[i], of_code = ofc[i], ml_code = mlc[i];
        int ll_bits = ZO_LL_BITS[ll_code], of_bits = of_code, ml_bits = ZO_ML_BITS[ml_code];
        of_state = fse_ct_encode(oft, &w, of_state, of_code);
        ml_state = fse_ct_encode(mlt, &w, ml_state, ml_code);
        ll_state = fse_ct_encode(llt, &w, ll_state, ll_code);
        if (of_bits + ml_bits + ll_bits >= 64 - 7 - (9 + 9 + 8)) bw_flush(&w);
        bw_add(&w, c->lit_lengths[i], ll_bits);
        if (ll_bits + ml_bits > 24) bw_flush(&w);
        bw_add(&w, c->match_lengths[i], ml_bits);
        if (of_bits + ml_bits + ll_bits > 56) bw_flush(&w);
        bw_add(&w, c->offsets[i], of_bits);
        bw_flush(&w);
    }
    fse_ct_finish(mlt, &w, ml_state);
    fse_ct_finish(oft, &w, of_state);
    fse_ct_finish(llt, &w, ll_state);
    int size = bw_close(&w);
    return size > 0 ? size : -1;
}

static int compress_sequences(zcctx *c, uint8_t *out, int64_t addr, int64_t out_size)   /* :66-209 */
{
    int64_t output = addr, limit = addr + out_size;
    if (!(limit - output > 3 + 1)) return -1;
    int n = c->seq_count;
    if (n < 0x7F) out[output++] = (uint8_t) n;
    else if (n < 0x7F00) { out[output] = (uint8_t) ((uint32_t) n >> 8 | 0x80); out[output + 1] = (uint8_t) n; output += 2; }
    else { out[output++] = 0xFF; zo_st16(out + output, (uint32_t) (n - 0x7F00)); output += 2; }
    if (n == 0) return (int) (output - addr);
    int64_t header = output++;
    int32_t counts[256];
    const fse_ctable *llt, *oft, *mlt;
    int r;

    histogram(c->ll_codes, n, counts, 256);
    int max_symbol = find_max_symbol(counts, 35);
    int largest = find_largest(counts, max_symbol);
    int ll_type = select_encoding(largest, n, 6, 1);
    if (ll_type == 1) { out[output++] = c->ll_codes[0]; fse_ct_rle(&c->ll_ct, max_symbol); llt = &c->ll_ct; }
    else if (ll_type == 0) llt = &c->def_ll;
    else { r = build_ctable(&c->ll_ct, out, output, limit, n, 9, c->ll_codes, counts, max_symbol); if (r < 0) return -1; output += r; llt = &c->ll_ct; }

    histogram(c->of_codes, n, counts, 256);
    max_symbol = find_max_symbol(counts, 31);
    largest = find_largest(counts, max_symbol);
    int of_type = select_encoding(largest, n, 5, max_symbol < 28);
    if (of_type == 1) { out[output++] = c->of_codes[0]; fse_ct_rle(&c->of_ct, max_symbol); oft = &c->of_ct; }
    else if (of_type == 0) oft = &c->def_of;
    else { r = build_ctable(&c->of_ct, out, output, limit, n, 8, c->of_codes, counts, max_symbol); if (r < 0) return -1; output += r; oft = &c->of_ct; }

    histogram(c->ml_codes, n, counts, 256);
    max_symbol = find_max_symbol(counts, 52);
    largest = find_largest(counts, max_symbol);
    int ml_type = select_encoding(largest, n, 6, 1);
    if (ml_type == 1) { out[output++] = c->ml_codes[0]; fse_ct_rle(&c->ml_ct, max_symbol); mlt = &c->ml_ct; }
    else if (ml_type == 0) mlt = &c->def_ml;
    else { r = build_ctable(&c->ml_ct, out, output, limit, n, 9, c->ml_codes, counts, max_symbol); if (r < 0) return -1; output += r; mlt = &c->ml_ct; }

    out[header] = (uint8_t) ((ll_type << 6) | (of_type << 4) | (ml_type << 2));
    r = encode_sequences(c, out, output, limit, mlt, oft, llt);
    if (r < 0) return -1;
    output += r;
    return (int) (output - addr);
}

/* ---- ZstdFrameCompressor.compressBlock :206-260 ----------------------------------------------------- */
static int compress_block(zcctx *c, const uint8_t *in, int64_t block_start, int32_t size, uint8_t *out, int64_t addr, int64_t out_size)
{
    if (size < 1 + 1 + 1 + 3 + 1) return 0;   /* MIN_BLOCK_SIZE + SIZE_OF_BLOCK_HEADER + 1 */
    /* enforceMaxDistance (BlockCompressionState.java:60-66) */
    int32_t distance = (int32_t) (block_start + size);
    int32_t new_offset = distance - (1 << c->p.window_log);
    if (c->window_base_offset < new_offset) c->window_base_offset = new_offset;
    c->literals_len = 0; c->seq_count = 0; c->long_field = 0;
    int32_t last_lits = dfast_compress_block(c, in, block_start, size);
    memcpy(c->literals + c->literals_len, in + block_start + size - last_lits, (size_t) last_lits);
    c->literals_len += last_lits;
    generate_codes(c);
    int64_t limit = addr + out_size, output = addr;
    int ls = encode_literals(c, out, output, limit - output, c->literals, c->literals_len);
    if (ls < 0) return -1;
    output += ls;
    int ss = compress_sequences(c, out, output, limit - output);
    if (ss < 0) return -1;
    int cs = ls + ss;
    if (cs == 0) return 0;
    if (cs > size - min_gain(size)) return 0;
    c->rep0 = c->tmp0; c->rep1 = c->tmp1;                        /* context.commit() */
    c->temp_table = c->temp_cand; c->prev_table = c->prev_cand;
    return cs;
}

/* ZstdFrameCompressor.compress :136-150 */
int64_t orc_zstd_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap)
{
    if (in_len > 0x7fffffff) return ORC_STATUS(ORC_E_ARGUMENT, 0);
    int32_t input_size = (int32_t) in_len;
    cparams p = compute_params(input_size);
    int64_t output = 0;
    if (out_cap - output < 4) ARG_FAIL();
    zo_st32(out + output, 0xFD2FB528u);                            /* writeMagic :52-58 */
    output += 4;
    if (out_cap - output < 6) ARG_FAIL();                          /* writeFrameHeader :61-120 (MAX_FRAME_HEADER_SIZE = 6) */
    int window_size = 1 << p.window_log;
    int cs_desc = (input_size >= 256) + (input_size >= 65536 + 256);
    int fhd = (cs_desc << 6) | 0x4;
    int single_segment = window_size >= input_size;
    if (single_segment) fhd |= 0x20;
    out[output++] = (uint8_t) fhd;
    if (!single_segment) {
        int base = 1 << zo_highbit((uint32_t) window_size);
        int exponent = zo_highbit((uint32_t) base);
        int remainder = window_size - base;
        int mantissa = remainder / (base / 8);
        out[output++] = (uint8_t) (((exponent - 10) << 3) | mantissa);
    }
    if (cs_desc == 0) { if (single_segment) out[output++] = (uint8_t) input_size; }
    else if (cs_desc == 1) { zo_st16(out + output, (uint32_t) (input_size - 256)); output += 2; }
    else { zo_st32(out + output, (uint32_t) input_size); output += 4; }

    /* compressFrame :152-179 */
    int block_size = window_size < ZO_MAX_BLOCK ? window_size : ZO_MAX_BLOCK;
    zcctx *c = (zcctx *) calloc(1, sizeof(zcctx));
    c->p = p;
    c->hash_table = (int32_t *) calloc((size_t) 1 << p.hash_log, 4);
    c->chain_table = (int32_t *) calloc((size_t) 1 << p.chain_log, 4);
    {   /* CompressionContext :31-44 */
        int ws = input_size < 1 ? 1 : (input_size > window_size ? window_size : input_size);
        int bs = ws < ZO_MAX_BLOCK ? ws : ZO_MAX_BLOCK;
        int max_seq = bs / (p.search_length == 3 ? 3 : 4);
        c->literals = (uint8_t *) malloc((size_t) bs + 16);
        c->offsets = (int32_t *) malloc(sizeof(int32_t) * (size_t) (max_seq + 1));
        c->lit_lengths = (int32_t *) malloc(sizeof(int32_t) * (size_t) (max_seq + 1));
        c->match_lengths = (int32_t *) malloc(sizeof(int32_t) * (size_t) (max_seq + 1));
        c->ll_codes = (uint8_t *) malloc((size_t) max_seq + 1);
        c->ml_codes = (uint8_t *) malloc((size_t) max_seq + 1);
        c->of_codes = (uint8_t *) malloc((size_t) max_seq + 1);
    }
    c->rep0 = 1; c->rep1 = 4;
    c->prev_table = c->prev_cand = &c->huf_a;
    c->temp_table = c->temp_cand = &c->huf_b;
    fse_ct_init(&c->def_ll, ZO_DEFAULT_LL_NORM, 35, 6);
    fse_ct_init(&c->def_ml, ZO_DEFAULT_ML_NORM, 52, 6);
    fse_ct_init(&c->def_of, ZO_DEFAULT_OF_NORM, 28, 5);

    int64_t result = 0;
    int64_t out_size = out_cap - output;
    int32_t remaining = input_size;
    int64_t input = 0;
    do {
        if (out_size < 3 + 3) { result = ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL); break; }
        int last_block = block_size >= remaining;
        if (block_size > remaining) block_size = remaining;
        /* writeCompressedBlock :181-204 */
        int cs = 0;
        if (block_size > 0) {
            cs = compress_block(c, in, input, block_size, out, output + 3, out_size - 3);
            if (cs < 0) { result = ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL); break; }
        }
        if (cs == 0) {
            if (block_size + 3 > out_size) { result = ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL); break; }
            zo_st24(out + output, (uint32_t) ((last_block ? 1 : 0) | (0 << 1) | (block_size << 3)));
            memcpy(out + output + 3, in + input, (size_t) block_size);
            cs = 3 + block_size;
        }
        else {
            zo_st24(out + output, (uint32_t) ((last_block ? 1 : 0) | (2 << 1) | (cs << 3)));
            cs += 3;
        }
        input += block_size;
        remaining -= block_size;
        output += cs;
        out_size -= cs;
    }
    while (remaining > 0);

    if (result == 0) {
        if (out_cap - output < 4) result = ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL);
        else { zo_st32(out + output, (uint32_t) orc_xxh64(in, in_len, 0)); output += 4; result = output; }   /* writeChecksum :123-134 */
    }
    free(c->hash_table); free(c->chain_table); free(c->literals); free(c->offsets); free(c->lit_lengths); free(c->match_lengths);
    free(c->ll_codes); free(c->ml_codes); free(c->of_codes); free(c);
    return result;
}
