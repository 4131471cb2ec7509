// zstd_fse_enc.cuh -- single-thread FSE encoder pieces for small alphabets (Huffman weights, sequence code tables).
// Follows FiniteStateEntropy.java:153-521 (optimalTableLog, normalizeCounts, normalizeCounts2, writeNormalizedCounts,
// compress) and HuffmanCompressionTable.compressWeights (:395-436) of the reference.
#pragma once
#include "zstd_common.cuh"

namespace zs {

__device__ inline int fse_min_table_log(int input_size, int max_symbol)
{
    int a = highbit((uint32_t) (input_size - 1)) + 1, b = highbit((uint32_t) max_symbol) + 2;
    return a < b ? a : b;
}

__device__ inline int fse_optimal_table_log(int max_table_log, int input_size, int max_symbol)
{
    int r = max_table_log, v = highbit((uint32_t) (input_size - 1)) - 2;
    if (v < r) r = v;
    v = fse_min_table_log(input_size, max_symbol);
    if (v > r) r = v;
    if (r < 5) r = 5;
    if (r > 12) r = 12;
    return r;
}

__device__ inline void fse_normalize2(int16_t *norm, int table_log, const int32_t *counts, int total, int max_symbol)
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
    const int factor = 1 << table_log;
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
        norm[max_value] = (int16_t) (norm[max_value] + to_distribute);
        return;
    }
    if (total == 0) {
        for (int i = 0; to_distribute > 0; i = (i + 1) % (max_symbol + 1)) {
            if (norm[i] > 0) { to_distribute--; norm[i]++; }
        }
        return;
    }
    const long long v_step_log = 62 - table_log;
    const long long mid = (1LL << (v_step_log - 1)) - 1;
    const long long r_step = (((1LL << v_step_log) * to_distribute) + mid) / total;
    long long tmp_total = mid;
    for (int i = 0; i <= max_symbol; i++) {
        if (norm[i] == UNASSIGNED) {
            long long end = tmp_total + ((long long) counts[i] * r_step);
            int s_start = (int) ((unsigned long long) tmp_total >> v_step_log);
            int s_end = (int) ((unsigned long long) end >> v_step_log);
            norm[i] = (int16_t) (s_end - s_start);
            tmp_total = end;
        }
    }
}

__device__ inline void fse_normalize(int16_t *norm, int table_log, const int32_t *counts, int total, int max_symbol)
{
    const int rest_to_beat[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    const long long scale = 62 - table_log;
    const long long step = (1LL << 62) / total;
    const long long vstep = 1LL << (scale - 20);
    int still = 1 << table_log;
    int largest = 0;
    int16_t largest_p = 0;
    const int low_threshold = (int) ((uint32_t) total >> table_log);
    for (int s = 0; s <= max_symbol; s++) {
        if (counts[s] == 0) { norm[s] = 0; continue; }
        if (counts[s] <= low_threshold) { norm[s] = -1; still--; }
        else {
            int16_t p = (int16_t) ((unsigned long long) ((long long) counts[s] * step) >> scale);
            if (p < 8) {
                long long rtb = vstep * rest_to_beat[p];
                long long delta = (long long) counts[s] * step - (((long long) p) << scale);
                if (delta > rtb) p++;
            }
            if (p > largest_p) { largest_p = p; largest = s; }
            norm[s] = p;
            still -= p;
        }
    }
    if (-still >= (int) ((uint32_t) (int32_t) norm[largest] >> 1)) fse_normalize2(norm, table_log, counts, total, max_symbol);
    else norm[largest] = (int16_t) (norm[largest] + still);
}

// writeNormalizedCounts (:407-521) into a byte buffer; returns size or -1
__device__ inline int fse_write_ncount(uint8_t *out, int cap, const int16_t *norm, int max_symbol, int table_log)
{
    int o = 0;
    const int table_size = 1 << table_log;
    int bit_count = 4;
    uint32_t bit_stream = (uint32_t) (table_log - 5);
    int remaining = table_size + 1, threshold = table_size, table_bits = table_log + 1;
    int symbol = 0;
    bool previous0 = false;
#define NC_FLUSH16() do { if (o + 2 > cap) return -1; out[o] = (uint8_t) bit_stream; out[o + 1] = (uint8_t) (bit_stream >> 8); o += 2; bit_stream >>= 16; } while (0)
    while (remaining > 1) {
        if (previous0) {
            int start = symbol;
            while (norm[symbol] == 0) symbol++;
            while (symbol >= start + 24) { start += 24; bit_stream |= 0xFFFFu << bit_count; NC_FLUSH16(); }
            while (symbol >= start + 3) { start += 3; bit_stream |= 3u << bit_count; bit_count += 2; }
            bit_stream |= (uint32_t) (symbol - start) << bit_count;
            bit_count += 2;
            if (bit_count > 16) { NC_FLUSH16(); bit_count -= 16; }
        }
        int count = norm[symbol++];
        const int max = (2 * threshold - 1) - remaining;
        remaining -= count < 0 ? -count : count;
        count++;
        if (count >= threshold) count += max;
        bit_stream |= (uint32_t) count << bit_count;
        bit_count += table_bits;
        bit_count -= (count < max ? 1 : 0);
        previous0 = (count == 1);
        while (remaining < threshold) { table_bits--; threshold >>= 1; }
        if (bit_count > 16) { NC_FLUSH16(); bit_count -= 16; }
    }
    if (o + 2 > cap) return -1;
    out[o] = (uint8_t) bit_stream; out[o + 1] = (uint8_t) (bit_stream >> 8);
    o += (bit_count + 7) / 8;
#undef NC_FLUSH16
    return o;
}

// small FSE encode table in thread-local arrays (tableLog <= 6, alphabet <= 13: Huffman weights)
struct SmallCTable {
    uint16_t next_state[64];
    int32_t dnb[16], dfs[16];
    int log2;
};

__device__ inline void small_ctable_build(SmallCTable &t, const int16_t *norm, int max_symbol, int table_log)
{
    uint8_t spread[64];
    int cumul[18];
    const int size = 1 << table_log;
    int high = size - 1;
    t.log2 = table_log;
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
    for (int i = 0; i < size; i++) { int s = spread[i]; t.next_state[cumul[s]++] = (uint16_t) (size + i); }
    int total = 0;
    for (int s = 0; s <= max_symbol; s++) {
        int n = norm[s];
        if (n == 0) t.dnb[s] = ((table_log + 1) << 16) - size;
        else if (n == -1 || n == 1) { t.dnb[s] = (table_log << 16) - size; t.dfs[s] = total - 1; total++; }
        else {
            int mbo = table_log - highbit((uint32_t) (n - 1));
            t.dnb[s] = (mbo << 16) - (n << mbo);
            t.dfs[s] = total - n;
            total += n;
        }
    }
}

struct ByteBitWriter {   // BitOutputStream.java:49-89 over a byte buffer
    uint8_t *out; int cap; int cur; uint64_t container; int bit_count;
    __device__ void add(uint32_t value, int bits) { container |= ((uint64_t) value & ((1ull << bits) - 1)) << bit_count; bit_count += bits; }
    __device__ void flush()
    {
        const int bytes = bit_count >> 3;
        for (int i = 0; i < 8 && cur + i < cap; i++) out[cur + i] = (uint8_t) (container >> (8 * i));
        cur += bytes;
        bit_count &= 7;
        container = bytes >= 8 ? 0 : container >> (bytes * 8);
    }
    __device__ int close() { add(1, 1); flush(); if (cur + 8 > cap) return 0; return cur + (bit_count > 0 ? 1 : 0); }
};

// HuffmanCompressionTable.compressWeights: returns the size of (FSE table description + bitstream) written to `out`,
// 0 when the weights are not compressible that way.
__device__ inline int fse_compress_weights(uint8_t *out, int cap, const uint8_t *weights, int n)
{
    if (n <= 1) return 0;
    int32_t counts[13];
    for (int i = 0; i < 13; i++) counts[i] = 0;
    for (int i = 0; i < n; i++) counts[weights[i]]++;
    int max_symbol = 12;
    while (counts[max_symbol] == 0) max_symbol--;
    int max_count = 0;
    for (int i = 0; i <= max_symbol; i++) if (counts[i] > max_count) max_count = counts[i];
    if (max_count == n || max_count == 1) return 0;
    int16_t norm[13];
    const int table_log = fse_optimal_table_log(6, n, max_symbol);
    fse_normalize(norm, table_log, counts, n, max_symbol);
    const int hs = fse_write_ncount(out, cap, norm, max_symbol, table_log);
    if (hs < 0) return 0;
    SmallCTable t;
    small_ctable_build(t, norm, max_symbol, table_log);
    if (cap - hs < 16) return 0;
    ByteBitWriter w;
    w.out = out + hs; w.cap = cap - hs; w.cur = 0; w.container = 0; w.bit_count = 0;
    auto begin = [&](int sym) { int ob = (int) ((uint32_t) (t.dnb[sym] + (1 << 15)) >> 16); int base = (int) ((uint32_t) ((ob << 16) - t.dnb[sym]) >> ob); return (int) t.next_state[base + t.dfs[sym]]; };
    auto encode = [&](int state, int sym) { int ob = (int) ((uint32_t) (state + t.dnb[sym]) >> 16); w.add((uint32_t) state, ob); return (int) t.next_state[(state >> ob) + t.dfs[sym]]; };
    // FiniteStateEntropy.compress :158-236
    int input = n, size = n;
    if (size <= 2) return 0;
    int state1, state2;
    if (size & 1) {
        state1 = begin(weights[--input]);
        state2 = begin(weights[--input]);
        state1 = encode(state1, weights[--input]);
        w.flush();
    }
    else {
        state2 = begin(weights[--input]);
        state1 = begin(weights[--input]);
    }
    size -= 2;
    if (size & 2) {
        state2 = encode(state2, weights[--input]);
        state1 = encode(state1, weights[--input]);
        w.flush();
    }
    while (input > 0) {
        state2 = encode(state2, weights[--input]);
        state1 = encode(state1, weights[--input]);
        state2 = encode(state2, weights[--input]);
        state1 = encode(state1, weights[--input]);
        w.flush();
    }
    w.add((uint32_t) state2, t.log2); w.flush();
    w.add((uint32_t) state1, t.log2); w.flush();
    const int cs = w.close();
    if (cs == 0) return 0;
    return hs + cs;
}

}  // namespace zs
