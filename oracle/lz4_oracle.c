/*
 * lz4_oracle.c -- CPU restatement of the reference's LZ4 block codec (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows  lz4/Lz4RawCompressor.java:50-311  and  lz4/Lz4RawDecompressor.java:35-198  of the reference.
 * Positions are indices into the caller's buffers instead of (base, absolute address) pairs.
 */
#include "oracle.h"
#include <string.h>

/* forward LZ77 copy with byte-copy semantics; 8 bytes at a time when the distance allows (the Java code
 * wild-copies longs too), never writing past dst+len */
static inline void orc_match_copy(uint8_t *dst, const uint8_t *src, int64_t len)
{
    int64_t dist = dst - src;
    if (dist >= 8) {
        while (len >= 8) { uint64_t v; memcpy(&v, src, 8); memcpy(dst, &v, 8); dst += 8; src += 8; len -= 8; }
    }
    while (len-- > 0) *dst++ = *src++;
}
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

enum { LAST_LITERALS = 5, MIN_MATCH = 4, MATCH_FIND_LIMIT = 12, MIN_LENGTH = 13, MAX_DISTANCE = 65535,
       HASH_LOG = 12, SKIP_TRIGGER = 6 };

/* Lz4RawCompressor.java:50-62 -- (value * 889523592379L >>> 28) & mask on Java longs */
static inline int32_t lz4_hash(uint64_t value, int32_t mask)
{
    return (int32_t) (((value * 889523592379ULL) >> 28) & (uint64_t) (uint32_t) mask);
}

/* Lz4RawCompressor.java:64-67 */
int64_t orc_lz4_max_compressed_length(int64_t n) { return n + n / 255 + 16; }

/* Lz4RawCompressor.java:304-311 -- smallest power of two > inputSize-1, clamped to [16, 4096] */
static int32_t lz4_table_size(int32_t input_size)
{
    uint32_t v = (uint32_t) (input_size - 1);
    uint32_t hob = 0;
    if (v != 0) { hob = 1u << (31 - __builtin_clz(v)); }
    int64_t target = (int32_t) (hob << 1); /* Integer.highestOneBit(x) << 1, int wrap */
    if (target < 16) target = 16;
    if (target > (1 << HASH_LOG)) target = 1 << HASH_LOG;
    return (int32_t) target;
}

/* Lz4RawCompressor.java:282-302 encodeRunLength */
static int64_t lz4_run_length(uint8_t *out, int64_t o, int64_t length)
{
    if (length >= 15) {
        out[o++] = 0xF0;
        int64_t remaining = length - 15;
        while (remaining >= 255) { out[o++] = 255; remaining -= 255; }
        out[o++] = (uint8_t) remaining;
    }
    else {
        out[o++] = (uint8_t) (length << 4);
    }
    return o;
}

/* Lz4RawCompressor.java:269-280 emitLastLiteral */
static int64_t lz4_last_literal(uint8_t *out, int64_t o, const uint8_t *in, int64_t from, int64_t length)
{
    o = lz4_run_length(out, o, length);
    memcpy(out + o, in + from, (size_t) length);
    return o + length;
}

/* Lz4RawCompressor.java:240-267 count -- bytes equal at (input, match), input bounded by limit */
static int32_t lz4_count(const uint8_t *in, int64_t input, int64_t limit, int64_t match)
{
    int32_t remaining = (int32_t) (limit - input);
    int32_t count = 0;
    while (count < remaining - 7) {
        uint64_t diff = ld64(in + match) ^ ld64(in + input);
        if (diff != 0) return count + (__builtin_ctzll(diff) >> 3);
        count += 8; input += 8; match += 8;
    }
    while (count < remaining && in[match] == in[input]) { count++; match++; input++; }
    return count;
}

/* Lz4RawCompressor.java:69-192 compress (table is the int[4096] owned by Lz4JavaCompressor) */
int64_t orc_lz4_compress(const uint8_t *in, int64_t in_len64, uint8_t *out, int64_t out_cap)
{
    if (in_len64 > 0x7E000000) return ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_INPUT_EXCEEDED);        /* :83-85 */
    if (out_cap < orc_lz4_max_compressed_length(in_len64)) return ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL); /* :87-89 */
    int32_t in_len = (int32_t) in_len64;
    int32_t table[1 << HASH_LOG];
    int32_t table_size = lz4_table_size(in_len);
    memset(table, 0, sizeof(int32_t) * (size_t) table_size);                                        /* :78-79 */
    int32_t mask = table_size - 1;

    int64_t input = 0, output = 0;
    const int64_t input_limit = in_len;
    const int64_t match_find_limit = input_limit - MATCH_FIND_LIMIT;
    const int64_t match_limit = input_limit - LAST_LITERALS;

    if (in_len < MIN_LENGTH) {                                                                      /* :98-101 */
        return lz4_last_literal(out, output, in, input, input_limit - input);
    }

    int64_t anchor = input;
    table[lz4_hash(ld64(in + input), mask)] = (int32_t) input;                                      /* :107 */
    input++;
    int32_t next_hash = lz4_hash(ld64(in + input), mask);

    int done = 0;
    do {
        int64_t next_input = input;
        int32_t attempts = 1 << SKIP_TRIGGER;
        int32_t step = 1;
        int64_t match;
        do {                                                                                        /* :119-138 */
            int32_t h = next_hash;
            input = next_input;
            next_input += step;
            step = (int32_t) ((uint32_t) (attempts++) >> SKIP_TRIGGER);
            if (next_input > match_find_limit) {
                return lz4_last_literal(out, output, in, anchor, input_limit - anchor);
            }
            match = table[h];
            next_hash = lz4_hash(ld64(in + next_input), mask);
            table[h] = (int32_t) input;
        }
        while (ld32(in + match) != ld32(in + input) || match + MAX_DISTANCE < input);

        while (input > anchor && match > 0 && in[input - 1] == in[match - 1]) { --input; --match; } /* :141-144 */

        int32_t literal_length = (int32_t) (input - anchor);
        int64_t token = output;
        /* emitLiteral :194-207 (the Java wild-copies 8 bytes at a time; the bytes kept are the same) */
        output = lz4_run_length(out, token, literal_length);
        memcpy(out + output, in + anchor, (size_t) literal_length);
        output += literal_length;

        for (;;) {                                                                                  /* :152-183 */
            int32_t match_length = lz4_count(in, input + MIN_MATCH, match_limit, match + MIN_MATCH);
            /* emitMatch :209-235 */
            uint16_t off = (uint16_t) (input - match);
            out[output] = (uint8_t) off; out[output + 1] = (uint8_t) (off >> 8);
            output += 2;
            if (match_length >= 15) {
                out[token] |= 15;
                int64_t remaining = match_length - 15;
                while (remaining >= 510) { out[output++] = 255; out[output++] = 255; remaining -= 510; }
                if (remaining >= 255) { out[output++] = 255; remaining -= 255; }
                out[output++] = (uint8_t) remaining;
            }
            else {
                out[token] |= (uint8_t) match_length;
            }
            input += match_length + MIN_MATCH;
            anchor = input;
            if (input > match_find_limit) { done = 1; break; }

            int64_t position = input - 2;
            table[lz4_hash(ld64(in + position), mask)] = (int32_t) position;

            int32_t h = lz4_hash(ld64(in + input), mask);
            match = table[h];
            table[h] = (int32_t) input;
            if (match + MAX_DISTANCE < input || ld32(in + match) != ld32(in + input)) {
                input++;
                next_hash = lz4_hash(ld64(in + input), mask);
                break;
            }
            token = output++;
            out[token] = 0;
        }
    }
    while (!done);

    return lz4_last_literal(out, output, in, anchor, input_limit - anchor);                         /* :189 */
}

/* Lz4RawDecompressor.java:35-198 decompress */
int64_t orc_lz4_decompress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, int64_t *err_offset)
{
#define FAIL(off, reason) do { if (err_offset) *err_offset = (off); return ORC_STATUS(ORC_E_MALFORMED, reason); } while (0)
    const int64_t fast_output_limit = out_cap - 8;
    int64_t input = 0, output = 0;

    if (in_len == 0) FAIL(0, ORC_R_INPUT_EMPTY);                                                    /* :48-50 */
    if (out_cap == 0) {                                                                             /* :52-57 */
        if (in_len == 1 && in[0] == 0) return 0;
        if (err_offset) *err_offset = 0;
        return ORC_STATUS(ORC_E_DST_TOO_SMALL, ORC_R_LZ4_ZERO_CAPACITY); /* Java returns -1 here */
    }

    while (input < in_len) {
        const int32_t token = in[input++];
        int32_t literal_length = token >> 4;
        if (literal_length == 0xF) {                                                                /* :63-74 */
            if (input >= in_len) FAIL(input, ORC_R_NONE);
            int32_t value;
            do {
                value = in[input++];
                literal_length = (int32_t) ((uint32_t) literal_length + (uint32_t) value);          /* Java int wrap */
            }
            while (value == 255 && input < in_len - 15);
        }
        if (literal_length < 0) FAIL(input, ORC_R_NONE);

        int64_t literal_end = input + literal_length;
        int64_t literal_output_limit = output + literal_length;
        if (literal_output_limit > (fast_output_limit - MIN_MATCH) || literal_end > in_len - (2 + 1 + LAST_LITERALS)) {   /* :82-96 */
            if (literal_output_limit > out_cap) FAIL(input, ORC_R_LAST_LITERAL_OUTSIDE);
            if (literal_end != in_len) FAIL(input, ORC_R_ALL_INPUT_CONSUMED);
            memcpy(out + output, in + input, (size_t) literal_length);
            output += literal_length;
            break;
        }
        memcpy(out + output, in + input, (size_t) literal_length);                                  /* :99-107 (wild copy) */
        output = literal_output_limit;
        input = literal_end;

        int32_t offset = in[input] | (in[input + 1] << 8);                                          /* :113-119 */
        input += 2;
        int64_t match = output - offset;
        if (match < 0 || match >= output) FAIL(input, ORC_R_OFFSET_OUTSIDE);

        int32_t match_length = token & 0xF;                                                         /* :122-138 */
        if (match_length == 0xF) {
            int32_t value;
            do {
                if (input > in_len - LAST_LITERALS) FAIL(input, ORC_R_NONE);
                value = in[input++];
                match_length = (int32_t) ((uint32_t) match_length + (uint32_t) value);
            }
            while (value == 255);
        }
        match_length = (int32_t) ((uint32_t) match_length + MIN_MATCH);
        if (match_length < 0) FAIL(input, ORC_R_NONE);

        int64_t match_output_limit = output + match_length;
        if (match_output_limit > fast_output_limit - MIN_MATCH) {                                   /* :168-171 */
            if (match_output_limit > out_cap - LAST_LITERALS) FAIL(input, ORC_R_LAST5_LITERALS);
        }
        /* :146-192 -- overlap-safe copy; the kept bytes equal a forward byte-by-byte copy */
        orc_match_copy(out + output, out + match, match_length);
        output = match_output_limit;
    }
    return output;
#undef FAIL
}
