/*
 * snappy_oracle.c -- CPU restatement of the reference's Snappy raw codec (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows  snappy/SnappyRawCompressor.java:47-411  and  snappy/SnappyRawDecompressor.java:30-321.
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
static inline uint16_t ld16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

enum { BLOCK_SIZE = 1 << 16, INPUT_MARGIN = 15, MAX_TABLE = 1 << 14 };

/* SnappyRawCompressor.java:47-70 */
int64_t orc_snappy_max_compressed_length(int64_t n) { return 32 + n + n / 6; }

/* SnappyRawCompressor.java:348-361 getHashTableSize: clamp(highestOneBit(n-1) << 1, 256, 16384) */
static int32_t snappy_table_size(int32_t input_size)
{
    uint32_t v = (uint32_t) (input_size - 1);
    uint32_t hob = v ? 1u << (31 - __builtin_clz(v)) : 0;
    int64_t target = (int32_t) (hob << 1);
    if (target < 256) target = 256;
    if (target > MAX_TABLE) target = MAX_TABLE;
    return (int32_t) target;
}

/* SnappyRawCompressor.java:368-371 hashBytes: (value * 0x1e35a7bd) >>> shift on Java ints */
static inline int32_t snappy_hash(uint32_t value, int32_t shift) { return (int32_t) ((value * 0x1e35a7bdu) >> shift); }

/* SnappyRawCompressor.java:235-266 count */
static int32_t snappy_count(const uint8_t *in, int64_t start, int64_t match, int64_t limit)
{
    int64_t cur = start;
    while (cur < limit - 7) {
        uint64_t diff = ld64(in + match) ^ ld64(in + cur);
        if (diff != 0) { cur += __builtin_ctzll(diff) >> 3; return (int32_t) (cur - start); }
        cur += 8; match += 8;
    }
    if (cur < limit - 3 && ld32(in + match) == ld32(in + cur)) { cur += 4; match += 4; }
    if (cur < limit - 1 && ld16(in + match) == ld16(in + cur)) { cur += 2; match += 2; }
    if (cur < limit && in[match] == in[cur]) ++cur;
    return (int32_t) (cur - start);
}

/* SnappyRawCompressor.java:268-298 emitLiteralLength */
static int64_t snappy_literal_tag(uint8_t *out, int64_t o, int32_t literal_length)
{
    int32_t n = literal_length - 1;
    if (n < 60) { out[o++] = (uint8_t) (n << 2); return o; }
    int bytes;
    if (n < (1 << 8)) { out[o++] = (uint8_t) ((59 + 1) << 2); bytes = 1; }
    else if (n < (1 << 16)) { out[o++] = (uint8_t) ((59 + 2) << 2); bytes = 2; }
    else if (n < (1 << 24)) { out[o++] = (uint8_t) ((59 + 3) << 2); bytes = 3; }
    else { out[o++] = (uint8_t) ((59 + 4) << 2); bytes = 4; }
    for (int i = 0; i < bytes; i++) out[o + i] = (uint8_t) ((uint32_t) n >> (8 * i));
    return o + bytes;
}

/* SnappyRawCompressor.java:312-345 emitCopy */
static int64_t snappy_emit_copy(uint8_t *out, int64_t o, int64_t offset, int32_t len)
{
    while (len >= 68) {
        out[o++] = (uint8_t) (2 + ((64 - 1) << 2)); out[o++] = (uint8_t) offset; out[o++] = (uint8_t) (offset >> 8);
        len -= 64;
    }
    if (len > 64) {
        out[o++] = (uint8_t) (2 + ((60 - 1) << 2)); out[o++] = (uint8_t) offset; out[o++] = (uint8_t) (offset >> 8);
        len -= 60;
    }
    if (len < 12 && offset < 2048) {
        out[o++] = (uint8_t) (1 + ((len - 4) << 2) + ((offset >> 8) << 5));
        out[o++] = (uint8_t) offset;
    }
    else {
        out[o++] = (uint8_t) (2 + ((len - 1) << 2)); out[o++] = (uint8_t) offset; out[o++] = (uint8_t) (offset >> 8);
    }
    return o;
}

/* SnappyRawCompressor.java:383-411 writeUncompressedLength */
static int64_t snappy_write_length(uint8_t *out, int64_t o, int32_t n)
{
    if (n < (1 << 7) && n >= 0) { out[o++] = (uint8_t) n; }
    else if (n < (1 << 14) && n > 0) { out[o++] = (uint8_t) (n | 0x80); out[o++] = (uint8_t) ((uint32_t) n >> 7); }
    else if (n < (1 << 21) && n > 0) {
        out[o++] = (uint8_t) (n | 0x80); out[o++] = (uint8_t) (((uint32_t) n >> 7) | 0x80); out[o++] = (uint8_t) ((uint32_t) n >> 14);
    }
    else if (n < (1 << 28) && n > 0) {
        out[o++] = (uint8_t) (n | 0x80); out[o++] = (uint8_t) (((uint32_t) n >> 7) | 0x80);
        out[o++] = (uint8_t) (((uint32_t) n >> 14) | 0x80); out[o++] = (uint8_t) ((uint32_t) n >> 21);
    }
    else {
        out[o++] = (uint8_t) (n | 0x80); out[o++] = (uint8_t) (((uint32_t) n >> 7) | 0x80);
        out[o++] = (uint8_t) (((uint32_t) n >> 14) | 0x80); out[o++] = (uint8_t) (((uint32_t) n >> 21) | 0x80);
        out[o++] = (uint8_t) ((uint32_t) n >> 28);
    }
    return o;
}

/* SnappyRawCompressor.java:74-233 compress */
int64_t orc_snappy_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap)
{
    if (out_cap < orc_snappy_max_compressed_length((int32_t) in_len)) return ORC_STATUS(ORC_E_ARGUMENT, ORC_R_MAX_OUTPUT_TOO_SMALL); /* :87-90 */
    uint16_t table[MAX_TABLE];
    int64_t output = snappy_write_length(out, 0, (int32_t) in_len);

    for (int64_t block = 0; block < in_len; block += BLOCK_SIZE) {                                  /* :93 */
        const int64_t block_limit = (in_len < block + BLOCK_SIZE) ? in_len : block + BLOCK_SIZE;
        int64_t input = block;
        int32_t table_size = snappy_table_size((int32_t) (block_limit - block));
        memset(table, 0, sizeof(uint16_t) * (size_t) table_size);                                   /* :98-99 */
        const int32_t shift = 32 - (31 - __builtin_clz((uint32_t) table_size));                     /* :102 */

        int64_t next_emit = input;
        const int64_t fast_limit = block_limit - INPUT_MARGIN;
        while (input <= fast_limit) {                                                               /* :111 */
            int32_t skip = 32;
            int64_t candidate = 0;
            for (input += 1; input + ((uint32_t) skip >> 5) <= fast_limit; input += ((uint32_t) (skip++) >> 5)) {   /* :142-158 */
                uint32_t cur = ld32(in + input);
                int32_t h = snappy_hash(cur, shift);
                candidate = block + table[h];
                table[h] = (uint16_t) (input - block);
                if (cur == ld32(in + candidate)) break;
            }
            if (input + ((uint32_t) skip >> 5) > fast_limit) break;                                 /* :160-162 */

            int32_t literal_length = (int32_t) (input - next_emit);                                 /* :169-175 */
            output = snappy_literal_tag(out, output, literal_length);
            memcpy(out + output, in + next_emit, (size_t) literal_length);                          /* fastCopy keeps exactly these bytes */
            output += literal_length;

            uint32_t input_bytes;
            do {                                                                                    /* :186-218 */
                int32_t matched = snappy_count(in, input + 4, candidate + 4, block_limit) + 4;
                output = snappy_emit_copy(out, output, input - candidate, matched);
                input += matched;
                if (input >= fast_limit) break;
                uint64_t lv = ld64(in + input - 1);
                uint32_t prev = (uint32_t) lv;
                input_bytes = (uint32_t) (lv >> 8);
                table[snappy_hash(prev, shift)] = (uint16_t) (input - block - 1);
                int32_t ch = snappy_hash(input_bytes, shift);
                candidate = block + table[ch];
                table[ch] = (uint16_t) (input - block);
            }
            while (input_bytes == ld32(in + candidate));
            next_emit = input;
        }
        if (next_emit < block_limit) {                                                              /* :224-229 */
            int32_t literal_length = (int32_t) (block_limit - next_emit);
            output = snappy_literal_tag(out, output, literal_length);
            memcpy(out + output, in + next_emit, (size_t) literal_length);
            output += literal_length;
        }
    }
    return output;
}

/* SnappyRawDecompressor.java:277-321 readUncompressedLength; returns length, *bytes_read out */
static int64_t snappy_read_length(const uint8_t *in, int64_t in_len, int32_t *bytes_read, int64_t *err_offset)
{
#define GETB(dst) do { if (n >= in_len) { if (err_offset) *err_offset = in_len - n; \
                       return ORC_STATUS(ORC_E_MALFORMED, ORC_R_SNAPPY_TRUNCATED); } dst = in[n]; n++; } while (0)
    int64_t n = 0;
    int32_t b;
    uint32_t result;
    GETB(b); result = (uint32_t) b & 0x7f;
    if (b & 0x80) {
        GETB(b); result |= ((uint32_t) b & 0x7f) << 7;
        if (b & 0x80) {
            GETB(b); result |= ((uint32_t) b & 0x7f) << 14;
            if (b & 0x80) {
                GETB(b); result |= ((uint32_t) b & 0x7f) << 21;
                if (b & 0x80) {
                    GETB(b); result |= ((uint32_t) b & 0x7f) << 28;
                    if (b & 0x80) {
                        if (err_offset) *err_offset = n; /* Java passes compressedAddress + bytesRead */
                        return ORC_STATUS(ORC_E_MALFORMED, ORC_R_SNAPPY_VARINT_HIGHBIT);
                    }
                }
            }
        }
    }
#undef GETB
    if ((int32_t) result < 0) { if (err_offset) *err_offset = 0; return ORC_STATUS(ORC_E_MALFORMED, ORC_R_SNAPPY_NEG_LENGTH); }
    *bytes_read = (int32_t) n;
    return (int32_t) result;
}

int64_t orc_snappy_uncompressed_length(const uint8_t *in, int64_t in_len, int64_t *err_offset)
{
    int32_t br;
    return snappy_read_length(in, in_len, &br, err_offset);
}

/* SnappyRawDecompressor.java:238-271 opLookupTable, restated as a formula:
 *   bits 0-7 length, bits 8-10 copy offset / 256, bits 11-13 trailer byte count */
static inline uint32_t snappy_op_entry(uint32_t op)
{
    uint32_t kind = op & 3, hi = op >> 2;
    if (kind == 0) return hi < 60 ? hi + 1 : (((hi - 59) << 11) | 1);
    if (kind == 1) return (1u << 11) | ((op >> 5) << 8) | (4 + (hi & 7));
    if (kind == 2) return (2u << 11) | (hi + 1);
    return (4u << 11) | (hi + 1);
}

/* SnappyRawDecompressor.java:35-220 decompress + uncompressAll */
int64_t orc_snappy_decompress(const uint8_t *in0, int64_t in_len0, uint8_t *out, int64_t out_cap, int64_t *err_offset)
{
#define FAIL(off) do { if (err_offset) *err_offset = (off); return ORC_STATUS(ORC_E_MALFORMED, ORC_R_NONE); } while (0)
    int32_t br = 0;
    int64_t expected = snappy_read_length(in0, in_len0, &br, err_offset);
    if (expected < 0) return expected;
    if (!(expected <= out_cap)) { if (err_offset) *err_offset = 0; return ORC_STATUS(ORC_E_DST_TOO_SMALL, ORC_R_SNAPPY_LEN_GT_CAP); } /* :49-50 */

    /* uncompressAll: offsets in exceptions are relative to the byte after the preamble (:70-76) */
    const uint8_t *in = in0 + br;
    const int64_t in_len = in_len0 - br;
    const int64_t fast_output_limit = out_cap - 8;
    int64_t input = 0, output = 0;
    static const uint32_t wordmask[5] = {0, 0xff, 0xffff, 0xffffff, 0xffffffff};

    while (input < in_len) {
        uint32_t op = in[input++];
        uint32_t entry = snappy_op_entry(op);
        int32_t trailer_bytes = (int32_t) (entry >> 11);
        int32_t trailer = 0;
        if (input + 4 < in_len) {                                                                   /* :89-91 */
            trailer = (int32_t) (ld32(in + input) & wordmask[trailer_bytes]);
        }
        else {
            if (input + trailer_bytes > in_len) FAIL(input);
            uint32_t t = 0;
            for (int i = trailer_bytes - 1; i >= 0; i--) t = (t << 8) | in[input + i];
            trailer = (int32_t) t;
        }
        if (trailer < 0) FAIL(input);
        input += trailer_bytes;

        int32_t length = (int32_t) (entry & 0xff);
        if (length == 0) continue;

        if ((op & 3) == 0) {
            int32_t literal_length = (int32_t) ((uint32_t) length + (uint32_t) trailer);
            if (literal_length < 0) FAIL(input);
            int64_t literal_output_limit = output + literal_length;
            if (literal_output_limit > fast_output_limit || input + literal_length > in_len - 8) {  /* :123-131 */
                if (literal_output_limit > out_cap || input + literal_length > in_len) FAIL(input);
            }
            memcpy(out + output, in + input, (size_t) literal_length);
            input += literal_length;
            output = literal_output_limit;
        }
        else {
            int32_t match_offset = (int32_t) (entry & 0x700);
            match_offset = (int32_t) ((uint32_t) match_offset + (uint32_t) trailer);
            if (match_offset <= 0) FAIL(input);                                                     /* :151-153 */
            int64_t match = output - match_offset;
            if (match < 0 || output + length > out_cap) FAIL(input);                                /* :156-158 */
            orc_match_copy(out + output, out + match, length);                  /* :164-214 forward-copy semantics */
            output += length;
        }
    }

    if (expected != output) { if (err_offset) *err_offset = 0; return ORC_STATUS(ORC_E_MALFORMED, ORC_R_SNAPPY_LEN_MISMATCH); } /* :61-65 */
    return expected;
#undef FAIL
}
