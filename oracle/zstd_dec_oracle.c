/*
 * zstd_dec_oracle.c -- CPU restatement of the reference's Zstandard frame decoder
 * (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows zstd/ZstdFrameDecompressor.java:135-962, zstd/Huffman.java:52-324,
 * zstd/FseTableReader.java:27-168, zstd/FiniteStateEntropy.java:38-151 (weights decode),
 * zstd/BitInputStream.java:34-205 and zstd/Constants.java.
 *
 * Differences that are deliberate and documented in DESIGN.md:
 *  - the Java exceptions carry absolute Unsafe addresses as "offset"; here offsets are relative to the
 *    start of the input buffer;
 *  - one call = one fresh ZstdFrameDecompressor (the Java object keeps its Huffman table across calls,
 *    ZstdFrameDecompressor.java:131; a call here starts with no table loaded);
 *  - the predefined FSE decode tables (ZstdFrameDecompressor.java:85-113) are rebuilt from the predefined
 *    distributions with the table builder instead of being restated as literals.
 */
#include "oracle.h"
#include "zstd_oracle_common.h"
#include <string.h>

#define FAIL(off, reason) do { if (err_offset) *err_offset = (off); return ORC_STATUS(ORC_E_MALFORMED, reason); } while (0)
#define CHECK(cond, off, reason) do { if (!(cond)) FAIL(off, reason); } while (0)
#define PROPAGATE(expr) do { int64_t r_ = (expr); if (r_ < 0) return r_; } while (0)

/* ---- BitInputStream.java ---------------------------------------------------------------------- */
typedef struct {
    const uint8_t *in;
    int64_t start, cur;
    uint64_t bits;
    int32_t consumed;
    int overflow;
} bitr;

/* BitInputStream.java:64-77 -- Java masks long shift counts to 6 bits */
static inline uint64_t peek_bits(int32_t consumed, uint64_t bits, int n) { return ((bits << (consumed & 63)) >> 1) >> ((63 - n) & 63); }
static inline uint64_t peek_bits_fast(int32_t consumed, uint64_t bits, int n) { return (bits << (consumed & 63)) >> ((64 - n) & 63); }

/* BitInputStream.Initializer.initialize :108-130 */
static int64_t br_init(bitr *b, const uint8_t *in, int64_t start, int64_t end, int64_t *err_offset)
{
    CHECK(end - start >= 1, start, ZR_BITSTREAM_EMPTY);
    int last = in[end - 1];
    CHECK(last != 0, end, ZR_BITSTREAM_NO_END_MARK);
    b->in = in; b->start = start; b->overflow = 0;
    b->consumed = 8 - zo_highbit((uint32_t) last);
    int64_t size = end - start;
    if (size >= 8) {
        b->cur = end - 8;
        b->bits = zo_ld64(in + b->cur);
    }
    else {
        b->cur = start;
        uint64_t v = in[start];
        for (int i = 1; i < size; i++) v |= (uint64_t) in[start + i] << (8 * i);   /* readTail :39-59 */
        b->bits = v;
        b->consumed += (int32_t) (8 - size) * 8;
    }
    return 0;
}

/* BitInputStream.Loader.load :171-204; returns the Java boolean ("done") */
static int br_load(bitr *b)
{
    if (b->consumed > 64) { b->overflow = 1; return 1; }
    if (b->cur == b->start) return 1;
    int32_t bytes = (int32_t) ((uint32_t) b->consumed >> 3);
    if (b->cur >= b->start + 8) {
        if (bytes > 0) { b->cur -= bytes; b->bits = zo_ld64(b->in + b->cur); }
        b->consumed &= 7;
    }
    else if (b->cur - bytes < b->start) {
        bytes = (int32_t) (b->cur - b->start);
        b->cur = b->start;
        b->consumed -= bytes * 8;
        b->bits = zo_ld64(b->in + b->start);
        return 1;
    }
    else {
        b->cur -= bytes;
        b->consumed -= bytes * 8;
        b->bits = zo_ld64(b->in + b->cur);
    }
    return 0;
}

/* ---- FSE decode tables: FiniteStateEntropy.Table :523-550 ------------------------------------- */
typedef struct {
    int log2;
    int32_t new_state[512];
    uint8_t symbol[512];
    uint8_t nbits[512];
} fse_dtable;

/* FseTableReader.java:111-159 -- populate a decode table from normalized counters */
static int fse_build_dtable(fse_dtable *t, const int16_t *norm, int max_symbol, int table_log)
{
    int16_t next[256];
    int size = 1 << table_log, high = size - 1;
    t->log2 = table_log;
    for (int s = 0; s <= max_symbol; s++) {
        if (norm[s] == -1) { t->symbol[high--] = (uint8_t) s; next[s] = 1; }
        else next[s] = norm[s];
    }
    int pos = zo_spread_symbols(norm, max_symbol, size, high, t->symbol);
    if (pos != 0) return -1;
    for (int i = 0; i < size; i++) {
        int s = t->symbol[i];
        int16_t ns = next[s]++;
        t->nbits[i] = (uint8_t) (table_log - zo_highbit((uint32_t) ns));
        t->new_state[i] = (int16_t) ((ns << t->nbits[i]) - size);
    }
    return 0;
}

/* FseTableReader.readFseTable :27-160; returns bytes consumed */
static int64_t fse_read_table(fse_dtable *t, const uint8_t *in, int64_t in_addr, int64_t in_limit, int max_symbol, int max_table_log,
                              int64_t *err_offset)
{
    int16_t norm[256];
    int64_t input = in_addr;
    CHECK(in_limit - in_addr >= 4, input, ZR_NOT_ENOUGH_INPUT);
    int symbol_number = 0, previous_is_zero = 0;
    uint32_t bit_stream = zo_ld32(in + input);
    int table_log = (int) (bit_stream & 0xF) + 5;
    int nbits = table_log + 1;
    bit_stream >>= 4;
    int bit_count = 4;
    CHECK(table_log <= max_table_log, input, ZR_FSE_TABLE_TOO_LARGE);
    int remaining = (1 << table_log) + 1;
    int threshold = 1 << table_log;

    while (remaining > 1 && symbol_number <= max_symbol) {
        if (previous_is_zero) {
            int n0 = symbol_number;
            while ((bit_stream & 0xFFFF) == 0xFFFF) {
                n0 += 24;
                if (input < in_limit - 5) { input += 2; bit_stream = zo_ld32(in + input) >> bit_count; }
                else { bit_stream >>= 16; bit_count += 16; }
            }
            while ((bit_stream & 3) == 3) { n0 += 3; bit_stream >>= 2; bit_count += 2; }
            n0 += (int) (bit_stream & 3);
            bit_count += 2;
            CHECK(n0 <= max_symbol, input, ZR_SYMBOL_TOO_LARGE);
            while (symbol_number < n0) norm[symbol_number++] = 0;
            if (input <= in_limit - 7 || input + (bit_count >> 3) <= in_limit - 4) {
                input += bit_count >> 3;
                bit_count &= 7;
                bit_stream = zo_ld32(in + input) >> bit_count;
            }
            else {
                bit_stream >>= 2;
            }
        }
        int16_t max = (int16_t) ((2 * threshold - 1) - remaining);
        int16_t count;
        if ((int32_t) (bit_stream & (uint32_t) (threshold - 1)) < max) {
            count = (int16_t) (bit_stream & (uint32_t) (threshold - 1));
            bit_count += nbits - 1;
        }
        else {
            count = (int16_t) (bit_stream & (uint32_t) (2 * threshold - 1));
            if (count >= threshold) count = (int16_t) (count - max);
            bit_count += nbits;
        }
        count--;
        remaining -= count < 0 ? -count : count;
        norm[symbol_number++] = count;
        previous_is_zero = count == 0;
        while (remaining < threshold) { nbits--; threshold >>= 1; }
        if (input <= in_limit - 7 || input + (bit_count >> 3) <= in_limit - 4) {
            input += bit_count >> 3;
            bit_count &= 7;
        }
        else {
            bit_count -= (int) (8 * (in_limit - 4 - input));
            input = in_limit - 4;
        }
        bit_stream = zo_ld32(in + input) >> (bit_count & 31);
    }
    CHECK(remaining == 1 && bit_count <= 32, input, ZR_CORRUPTED);
    int max_sym = symbol_number - 1;
    CHECK(max_sym <= 255, input, ZR_TOO_MANY_SYMBOLS);
    input += (bit_count + 7) >> 3;
    if (fse_build_dtable(t, norm, max_sym, table_log) != 0) FAIL(input, ZR_CORRUPTED);
    return input - in_addr;
}

/* FseTableReader.initializeRleTable :162-168 */
static void fse_rle_table(fse_dtable *t, uint8_t value)
{
    t->log2 = 0; t->symbol[0] = value; t->new_state[0] = 0; t->nbits[0] = 0;
}

/* FiniteStateEntropy.decompress :38-151 (Huffman weights); returns number of symbols */
static int64_t fse_decompress(const fse_dtable *t, const uint8_t *in, int64_t in_addr, int64_t in_limit, uint8_t *out, int out_cap,
                              int64_t *err_offset)
{
    bitr b;
    int64_t input = in_addr;
    int output = 0;
    PROPAGATE(br_init(&b, in, input, in_limit, err_offset));
    int state1 = (int) peek_bits(b.consumed, b.bits, t->log2); b.consumed += t->log2;
    br_load(&b);
    int state2 = (int) peek_bits(b.consumed, b.bits, t->log2); b.consumed += t->log2;
    br_load(&b);
#define FSE_STEP(st) do { int nb_ = t->nbits[st]; st = (int) (t->new_state[st] + (int64_t) peek_bits(b.consumed, b.bits, nb_)); b.consumed += nb_; } while (0)
    while (output <= out_cap - 4) {
        out[output] = t->symbol[state1]; FSE_STEP(state1);
        out[output + 1] = t->symbol[state2]; FSE_STEP(state2);
        out[output + 2] = t->symbol[state1]; FSE_STEP(state1);
        out[output + 3] = t->symbol[state2]; FSE_STEP(state2);
        output += 4;
        if (br_load(&b)) break;
    }
    for (;;) {
        CHECK(output <= out_cap - 2, input, ZR_FSE_OUTPUT_TOO_SMALL);
        out[output++] = t->symbol[state1]; FSE_STEP(state1);
        br_load(&b);
        if (b.overflow) { out[output++] = t->symbol[state2]; break; }
        CHECK(output <= out_cap - 2, input, ZR_FSE_OUTPUT_TOO_SMALL);
        out[output++] = t->symbol[state2]; FSE_STEP(state2);
        br_load(&b);
        if (b.overflow) { out[output++] = t->symbol[state1]; break; }
    }
#undef FSE_STEP
    return output;
}

/* ---- Huffman.java ----------------------------------------------------------------------------- */
typedef struct {
    int table_log;  /* -1 = not loaded */
    uint8_t symbols[4096];
    uint8_t nbits[4096];
} huf_dtable;

/* Huffman.readTable :52-128; returns bytes consumed */
static int64_t huf_read_table(huf_dtable *h, const uint8_t *in, int64_t in_addr, int size, int64_t *err_offset)
{
    uint8_t weights[257];
    int ranks[13 + 1];
    memset(ranks, 0, sizeof(ranks));
    memset(weights, 0, sizeof(weights));
    int64_t input = in_addr;
    CHECK(size > 0, input, ZR_NOT_ENOUGH_INPUT);
    int input_size = in[input++];
    int output_size;
    if (input_size >= 128) {
        output_size = input_size - 127;
        input_size = (output_size + 1) / 2;
        CHECK(input_size + 1 <= size, input, ZR_NOT_ENOUGH_INPUT);
        CHECK(output_size <= 256, input, ZR_CORRUPTED);
        for (int i = 0; i < output_size; i += 2) {
            int v = in[input + i / 2];
            weights[i] = (uint8_t) (v >> 4);
            weights[i + 1] = (uint8_t) (v & 15);
        }
    }
    else {
        CHECK(input_size + 1 <= size, input, ZR_NOT_ENOUGH_INPUT);
        int64_t limit = input + input_size;
        fse_dtable t;
        int64_t used = fse_read_table(&t, in, input, limit, 255, 6, err_offset);
        if (used < 0) return used;
        input += used;
        int64_t n = fse_decompress(&t, in, input, limit, weights, 256, err_offset);
        if (n < 0) return n;
        output_size = (int) n;
    }
    int total_weight = 0;
    for (int i = 0; i < output_size; i++) {
        CHECK(weights[i] <= 12, input, ZR_CORRUPTED);   /* Java would throw ArrayIndexOutOfBounds on ranks[] */
        ranks[weights[i]]++;
        total_weight += (1 << weights[i]) >> 1;
    }
    CHECK(total_weight != 0, input, ZR_CORRUPTED);
    int table_log = zo_highbit((uint32_t) total_weight) + 1;
    CHECK(table_log <= 12, input, ZR_CORRUPTED);
    int total = 1 << table_log;
    int rest = total - total_weight;
    CHECK((rest & (rest - 1)) == 0, input, ZR_CORRUPTED);
    int last_weight = zo_highbit((uint32_t) rest) + 1;
    weights[output_size] = (uint8_t) last_weight;
    ranks[last_weight]++;
    int number_of_symbols = output_size + 1;

    int next_rank_start = 0;
    for (int i = 1; i < table_log + 1; ++i) {
        int current = next_rank_start;
        next_rank_start += ranks[i] << (i - 1);
        ranks[i] = current;
    }
    for (int n = 0; n < number_of_symbols; n++) {
        int weight = weights[n];
        int length = (1 << weight) >> 1;
        uint8_t nb = (uint8_t) (table_log + 1 - weight);
        for (int i = ranks[weight]; i < ranks[weight] + length; i++) { h->symbols[i] = (uint8_t) n; h->nbits[i] = nb; }
        ranks[weight] += length;
    }
    CHECK(ranks[1] >= 2 && (ranks[1] & 1) == 0, input, ZR_CORRUPTED);
    h->table_log = table_log;
    return input_size + 1;
}

#define HUF_SYM(b, dst) do { int v_ = (int) peek_bits_fast((b).consumed, (b).bits, tl); *(dst) = h->symbols[v_]; (b).consumed += h->nbits[v_]; } while (0)

/* Huffman.decodeTail :291-317 */
static int64_t huf_decode_tail(const huf_dtable *h, bitr *b, uint8_t *out, int64_t o, int64_t o_limit, int64_t *err_offset)
{
    const int tl = h->table_log;
    while (o < o_limit) {
        if (br_load(b)) break;
        HUF_SYM(*b, out + o); o++;
    }
    while (o < o_limit) { HUF_SYM(*b, out + o); o++; }
    CHECK(b->start == b->cur && b->consumed == 64, b->start, ZR_BITSTREAM_NOT_CONSUMED);
    return 0;
}

/* Huffman.decodeSingleStream :130-164 */
static int64_t huf_decode_1(const huf_dtable *h, const uint8_t *in, int64_t in_addr, int64_t in_limit, uint8_t *out, int64_t o_limit,
                            int64_t *err_offset)
{
    const int tl = h->table_log;
    bitr b;
    PROPAGATE(br_init(&b, in, in_addr, in_limit, err_offset));
    int64_t o = 0;
    int64_t fast = o_limit - 4;
    while (o < fast) {
        if (br_load(&b)) break;
        HUF_SYM(b, out + o); HUF_SYM(b, out + o + 1); HUF_SYM(b, out + o + 2); HUF_SYM(b, out + o + 3);
        o += 4;
    }
    return huf_decode_tail(h, &b, out, o, o_limit, err_offset);
}

/* Huffman.decode4Streams :166-289 */
static int64_t huf_decode_4(const huf_dtable *h, const uint8_t *in, int64_t in_addr, int64_t in_limit, uint8_t *out, int64_t o_limit,
                            int64_t *err_offset)
{
    const int tl = h->table_log;
    CHECK(in_limit - in_addr >= 10, in_addr, ZR_CORRUPTED);
    int64_t start1 = in_addr + 6;
    int64_t start2 = start1 + zo_ld16(in + in_addr);
    int64_t start3 = start2 + zo_ld16(in + in_addr + 2);
    int64_t start4 = start3 + zo_ld16(in + in_addr + 4);
    CHECK(start2 < start3 && start3 < start4 && start4 < in_limit, in_addr, ZR_CORRUPTED);
    bitr b1, b2, b3, b4;
    PROPAGATE(br_init(&b1, in, start1, start2, err_offset));
    PROPAGATE(br_init(&b2, in, start2, start3, err_offset));
    PROPAGATE(br_init(&b3, in, start3, start4, err_offset));
    PROPAGATE(br_init(&b4, in, start4, in_limit, err_offset));
    int64_t seg = (o_limit + 3) / 4;
    int64_t os2 = seg, os3 = os2 + seg, os4 = os3 + seg;
    int64_t o1 = 0, o2 = os2, o3 = os3, o4 = os4;
    int64_t fast = o_limit - 7;
    while (o4 < fast) {
        for (int k = 0; k < 4; k++) {
            HUF_SYM(b1, out + o1 + k); HUF_SYM(b2, out + o2 + k); HUF_SYM(b3, out + o3 + k); HUF_SYM(b4, out + o4 + k);
        }
        o1 += 4; o2 += 4; o3 += 4; o4 += 4;
        if (br_load(&b1)) break;
        if (br_load(&b2)) break;
        if (br_load(&b3)) break;
        if (br_load(&b4)) break;
    }
    CHECK(o1 <= os2 && o2 <= os3 && o3 <= os4, in_addr, ZR_CORRUPTED);
    PROPAGATE(huf_decode_tail(h, &b1, out, o1, os2, err_offset));
    PROPAGATE(huf_decode_tail(h, &b2, out, o2, os3, err_offset));
    PROPAGATE(huf_decode_tail(h, &b3, out, o3, os4, err_offset));
    PROPAGATE(huf_decode_tail(h, &b4, out, o4, o_limit, err_offset));
    return 0;
}

/* ---- frame decoder state ---------------------------------------------------------------------- */
typedef struct {
    uint8_t literals[ZO_MAX_BLOCK + 8 + 32];
    const uint8_t *lit;      /* current literals */
    int64_t lit_size;
    int32_t prev[3];
    fse_dtable ll, of, ml;
    fse_dtable def_ll, def_of, def_ml;
    const fse_dtable *cur_ll, *cur_of, *cur_ml;
    huf_dtable huf;
} zdctx;

static const int32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64,
                                    0x80, 0x100, 0x200, 0x400, 0x800, 0x1000, 0x2000, 0x4000, 0x8000, 0x10000};
static const int32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30,
                                    31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403, 0x803, 0x1003,
                                    0x2003, 0x4003, 0x8003, 0x10003};

/* ZstdFrameDecompressor.readFrameHeader :860-940 */
typedef struct { int64_t header_size; int32_t window_size; int64_t content_size; int has_checksum; } frame_header;

static int64_t read_frame_header(frame_header *fh, const uint8_t *in, int64_t in_addr, int64_t in_limit, int64_t *err_offset)
{
    int64_t input = in_addr;
    CHECK(input < in_limit, input, ZR_NOT_ENOUGH_INPUT);
    int fhd = in[input++];
    int single_segment = (fhd & 0x20) != 0;
    int dict_desc = fhd & 3;
    int cs_desc = fhd >> 6;
    int header_size = 1 + (single_segment ? 0 : 1) + (dict_desc == 0 ? 0 : (1 << (dict_desc - 1))) +
                      (cs_desc == 0 ? (single_segment ? 1 : 0) : (1 << cs_desc));
    CHECK(header_size <= in_limit - in_addr, input, ZR_NOT_ENOUGH_INPUT);
    fh->window_size = -1;
    if (!single_segment) {
        int wd = in[input++];
        int exponent = wd >> 3, mantissa = wd & 7;
        int32_t base = (int32_t) (1u << ((10 + exponent) & 31));   /* Java int shift masks the count to 5 bits */
        fh->window_size = (int32_t) ((uint32_t) base + (uint32_t) (base / 8) * (uint32_t) mantissa);
    }
    int64_t dict_id = -1;
    if (dict_desc == 1) { dict_id = in[input]; input += 1; }
    else if (dict_desc == 2) { dict_id = zo_ld16(in + input); input += 2; }
    else if (dict_desc == 3) { dict_id = zo_ld32(in + input); input += 4; }
    CHECK(dict_id == -1, input, ZR_DICTIONARY);
    fh->content_size = -1;
    switch (cs_desc) {
        case 0: if (single_segment) { fh->content_size = in[input]; input += 1; } break;
        case 1: fh->content_size = (int64_t) zo_ld16(in + input) + 256; input += 2; break;
        case 2: fh->content_size = zo_ld32(in + input); input += 4; break;
        default: fh->content_size = (int64_t) zo_ld64(in + input); input += 8; break;
    }
    fh->has_checksum = (fhd & 4) != 0;
    fh->header_size = input - in_addr;
    return 0;
}

/* ZstdFrameDecompressor.verifyMagic :949-962 */
static int64_t verify_magic(const uint8_t *in, int64_t in_addr, int64_t in_limit, int64_t *err_offset)
{
    CHECK(in_limit - in_addr >= 4, in_addr, ZR_NOT_ENOUGH_INPUT);
    uint32_t magic = zo_ld32(in + in_addr);
    if (magic != 0xFD2FB528u) {
        if (magic == 0xFD2FB527u) FAIL(in_addr, ZR_V07_FORMAT);
        FAIL(in_addr, ZR_BAD_MAGIC);
    }
    return 4;
}

int64_t orc_zstd_decompressed_size(const uint8_t *in, int64_t in_len, int64_t *err_offset)
{
    /* ZstdFrameDecompressor.getDecompressedSize :942-947 */
    int64_t m = verify_magic(in, 0, in_len, err_offset);
    if (m < 0) return m;
    frame_header fh;
    PROPAGATE(read_frame_header(&fh, in, m, in_len, err_offset));
    return fh.content_size;  /* -1 when the frame does not record it */
}

/* decodeRawLiterals :812-858 */
static int64_t decode_raw_literals(zdctx *c, const uint8_t *in, int64_t in_addr, int64_t in_limit, int64_t *err_offset)
{
    int64_t input = in_addr;
    int type = (in[input] >> 2) & 3;
    int32_t lit_size;
    if (type == 0 || type == 2) { lit_size = in[input] >> 3; input++; }
    else if (type == 1) { lit_size = zo_ld16(in + input) >> 4; input += 2; }
    else { lit_size = (int32_t) ((in[input] | (zo_ld16(in + input + 1) << 8)) >> 4); input += 3; }
    CHECK(input + lit_size <= in_limit, input, ZR_NOT_ENOUGH_INPUT);
    c->lit = in + input;   /* same bytes whether the Java aliases the input or copies (:842-854) */
    c->lit_size = lit_size;
    input += lit_size;
    return input - in_addr;
}

/* decodeRleLiterals :776-810 */
static int64_t decode_rle_literals(zdctx *c, const uint8_t *in, int64_t in_addr, int block_size, int64_t *err_offset)
{
    int64_t input = in_addr;
    int32_t out_size;
    int type = (in[input] >> 2) & 3;
    if (type == 0 || type == 2) { out_size = in[input] >> 3; input++; }
    else if (type == 1) { out_size = zo_ld16(in + input) >> 4; input += 2; }
    else {
        CHECK(block_size >= 4, input, ZR_NOT_ENOUGH_INPUT);
        out_size = (int32_t) ((zo_ld32(in + input) & 0xFFFFFF) >> 4);
        input += 3;
    }
    CHECK(out_size <= ZO_MAX_BLOCK, input, ZR_OUTPUT_EXCEEDS_BLOCK);
    uint8_t value = in[input++];
    memset(c->literals, value, (size_t) out_size + 8);
    c->lit = c->literals;
    c->lit_size = out_size;
    return input - in_addr;
}

/* decodeCompressedLiterals :708-774 */
static int64_t decode_compressed_literals(zdctx *c, const uint8_t *in, int64_t in_addr, int block_size, int lit_type, int64_t *err_offset)
{
    int64_t input = in_addr;
    CHECK(block_size >= 5, input, ZR_NOT_ENOUGH_INPUT);
    int32_t comp_size, unc_size, header_size;
    int single = 0;
    int type = (in[input] >> 2) & 3;
    if (type == 0 || type == 1) {
        single = type == 0;
        uint32_t hd = zo_ld32(in + input);
        header_size = 3; unc_size = (int32_t) ((hd >> 4) & 0x3FF); comp_size = (int32_t) ((hd >> 14) & 0x3FF);
    }
    else if (type == 2) {
        uint32_t hd = zo_ld32(in + input);
        header_size = 4; unc_size = (int32_t) ((hd >> 4) & 0x3FFF); comp_size = (int32_t) ((hd >> 18) & 0x3FFF);
    }
    else {
        uint64_t hd = (uint64_t) in[input] | ((uint64_t) zo_ld32(in + input + 1) << 8);
        header_size = 5; unc_size = (int32_t) ((hd >> 4) & 0x3FFFF); comp_size = (int32_t) ((hd >> 22) & 0x3FFFF);
    }
    CHECK(unc_size <= ZO_MAX_BLOCK, input, ZR_BLOCK_EXCEEDS_MAX);
    CHECK(header_size + comp_size <= block_size, input, ZR_CORRUPTED);
    input += header_size;
    int64_t limit = input + comp_size;
    if (lit_type != 3) {
        int64_t used = huf_read_table(&c->huf, in, input, comp_size, err_offset);
        if (used < 0) return used;
        input += used;
    }
    c->lit = c->literals;
    c->lit_size = unc_size;
    if (single) PROPAGATE(huf_decode_1(&c->huf, in, input, limit, c->literals, unc_size, err_offset));
    else PROPAGATE(huf_decode_4(&c->huf, in, input, limit, c->literals, unc_size, err_offset));
    return header_size + comp_size;
}

/* computeLiteralsTable / computeOffsetsTable / computeMatchLengthTable :609-676 */
static int64_t compute_table(int type, fse_dtable *work, const fse_dtable *def, const fse_dtable **cur, int max_sym, int max_log,
                             const uint8_t *in, int64_t input, int64_t in_limit, int64_t *err_offset)
{
    switch (type) {
        case 1: {
            CHECK(input < in_limit, input, ZR_NOT_ENOUGH_INPUT);
            int8_t value = (int8_t) in[input++];
            CHECK(value <= max_sym, input, ZR_VALUE_EXCEEDS_MAX);
            CHECK(value >= 0, input, ZR_VALUE_EXCEEDS_MAX);   /* the Java lets a negative byte through and then fails with an array index error */
            fse_rle_table(work, (uint8_t) value);
            *cur = work;
            break;
        }
        case 0: *cur = def; break;
        case 3: CHECK(*cur != 0, input, ZR_EXPECTED_TABLE); break;
        default: {
            int64_t used = fse_read_table(work, in, input, in_limit, max_sym, max_log, err_offset);
            if (used < 0) return used;
            input += used;
            *cur = work;
        }
    }
    return input;
}

/* decompressSequences :312-516; returns bytes produced by this block */
static int64_t decompress_sequences(zdctx *c, const uint8_t *in, int64_t in_addr, int64_t in_limit, uint8_t *out, int64_t out_addr,
                                    int64_t out_limit, int64_t *err_offset)
{
    int64_t input = in_addr, output = out_addr;
    int64_t lit_pos = 0;
    CHECK(in_limit - in_addr >= 1, input, ZR_NOT_ENOUGH_INPUT);
    int32_t seq_count = in[input++];
    if (seq_count != 0) {
        if (seq_count == 255) {
            CHECK(input + 2 <= in_limit, input, ZR_NOT_ENOUGH_INPUT);
            seq_count = zo_ld16(in + input) + 0x7F00;
            input += 2;
        }
        else if (seq_count > 127) {
            CHECK(input < in_limit, input, ZR_NOT_ENOUGH_INPUT);
            seq_count = ((seq_count - 128) << 8) + in[input++];
        }
        CHECK(input + 4 <= in_limit, input, ZR_NOT_ENOUGH_INPUT);
        uint8_t type = in[input++];
        int ll_type = type >> 6, of_type = (type >> 4) & 3, ml_type = (type >> 2) & 3;
        int64_t r;
        r = compute_table(ll_type, &c->ll, &c->def_ll, &c->cur_ll, 35, 9, in, input, in_limit, err_offset); if (r < 0) return r; input = r;
        r = compute_table(of_type, &c->of, &c->def_of, &c->cur_of, 28, 8, in, input, in_limit, err_offset); if (r < 0) return r; input = r;
        r = compute_table(ml_type, &c->ml, &c->def_ml, &c->cur_ml, 52, 9, in, input, in_limit, err_offset); if (r < 0) return r; input = r;

        bitr b;
        PROPAGATE(br_init(&b, in, input, in_limit, err_offset));
        const fse_dtable *tl = c->cur_ll, *to = c->cur_of, *tm = c->cur_ml;
        int ll_state = (int) peek_bits(b.consumed, b.bits, tl->log2); b.consumed += tl->log2;
        int of_state = (int) peek_bits(b.consumed, b.bits, to->log2); b.consumed += to->log2;
        int ml_state = (int) peek_bits(b.consumed, b.bits, tm->log2); b.consumed += tm->log2;
        int32_t *prev = c->prev;

        while (seq_count > 0) {
            seq_count--;
            br_load(&b);
            if (b.overflow) {
                CHECK(seq_count == 0, input, ZR_NOT_ALL_SEQUENCES);
                break;
            }
            int ll_code = tl->symbol[ll_state];
            int ml_code = tm->symbol[ml_state];
            int of_code = to->symbol[of_state];
            /* the Java indexes fixed-size arrays with these codes; out-of-range codes cannot come out of
             * tables validated above (max symbol 35 / 52 / 28) except through predefined/RLE paths, also bounded */
            int ll_bits = ZO_LL_BITS[ll_code], ml_bits = ZO_ML_BITS[ml_code], of_bits = of_code;
            int32_t offset = ZO_OF_BASE[of_code];
            if (of_code > 0) { offset += (int32_t) peek_bits(b.consumed, b.bits, of_bits); b.consumed += of_bits; }
            if (of_code <= 1) {
                if (ll_code == 0) offset++;
                if (offset != 0) {
                    int32_t temp = (offset == 3) ? prev[0] - 1 : prev[offset];
                    if (temp == 0) temp = 1;
                    if (offset != 1) prev[2] = prev[1];
                    prev[1] = prev[0];
                    prev[0] = temp;
                    offset = temp;
                }
                else {
                    offset = prev[0];
                }
            }
            else {
                prev[2] = prev[1]; prev[1] = prev[0]; prev[0] = offset;
            }
            int32_t match_length = ML_BASE[ml_code];
            if (ml_code > 31) { match_length += (int32_t) peek_bits(b.consumed, b.bits, ml_bits); b.consumed += ml_bits; }
            int32_t lit_length = LL_BASE[ll_code];
            if (ll_code > 15) { lit_length += (int32_t) peek_bits(b.consumed, b.bits, ll_bits); b.consumed += ll_bits; }
            int total_bits = ll_bits + ml_bits + of_bits;
            if (total_bits > 64 - 7 - (9 + 9 + 8)) br_load(&b);

            int nb;
            nb = tl->nbits[ll_state]; ll_state = (int) (tl->new_state[ll_state] + (int64_t) peek_bits(b.consumed, b.bits, nb)); b.consumed += nb;
            nb = tm->nbits[ml_state]; ml_state = (int) (tm->new_state[ml_state] + (int64_t) peek_bits(b.consumed, b.bits, nb)); b.consumed += nb;
            nb = to->nbits[of_state]; of_state = (int) (to->new_state[of_state] + (int64_t) peek_bits(b.consumed, b.bits, nb)); b.consumed += nb;

            int64_t lit_out_limit = output + lit_length;
            int64_t match_out_limit = lit_out_limit + match_length;
            CHECK(match_out_limit <= out_limit, input, ZR_OUTPUT_TOO_SMALL);
            int64_t lit_end = lit_pos + lit_length;
            CHECK(lit_end <= c->lit_size, input, ZR_CORRUPTED);
            int64_t match = lit_out_limit - offset;
            CHECK(match >= 0, input, ZR_CORRUPTED);   /* outputAbsoluteBaseAddress = start of the caller's buffer */
            memcpy(out + output, c->lit + lit_pos, (size_t) lit_length);
            zo_match_copy(out + lit_out_limit, out + match, match_length);
            output = match_out_limit;
            lit_pos = lit_end;
        }
    }
    /* copyLastLiteral :518-525 */
    int64_t last = c->lit_size - lit_pos;
    CHECK(output + last <= out_limit, input, ZR_OUTPUT_TOO_SMALL);
    memcpy(out + output, c->lit + lit_pos, (size_t) last);
    output += last;
    return output - out_addr;
}

/* decodeCompressedBlock :265-310 */
static int64_t decode_compressed_block(zdctx *c, const uint8_t *in, int64_t in_addr, int block_size, uint8_t *out, int64_t out_addr,
                                       int64_t out_limit, int32_t window_size, int64_t *err_offset)
{
    int64_t input = in_addr;
    CHECK(block_size <= ZO_MAX_BLOCK, input, ZR_EXPECTED_TABLE);   /* message reused by the Java, :278 */
    CHECK(block_size >= 3, input, ZR_BLOCK_TOO_SMALL);
    int lit_type = in[input] & 3;
    int64_t used;
    if (lit_type == 0) used = decode_raw_literals(c, in, input, in_addr + block_size, err_offset);
    else if (lit_type == 1) used = decode_rle_literals(c, in, input, block_size, err_offset);
    else {
        if (lit_type == 3) CHECK(c->huf.table_log != -1, input, ZR_DICTIONARY_CORRUPTED);
        used = decode_compressed_literals(c, in, input, block_size, lit_type, err_offset);
    }
    if (used < 0) return used;
    input += used;
    CHECK(window_size <= (1 << 23), input, ZR_WINDOW_TOO_LARGE);
    return decompress_sequences(c, in, input, in_addr + block_size, out, out_addr, out_limit, err_offset);
}

static void zdctx_init(zdctx *c)
{
    c->huf.table_log = -1;
    fse_build_dtable(&c->def_ll, ZO_DEFAULT_LL_NORM, 35, 6);
    fse_build_dtable(&c->def_of, ZO_DEFAULT_OF_NORM, 28, 5);
    fse_build_dtable(&c->def_ml, ZO_DEFAULT_ML_NORM, 52, 6);
}

/* ZstdFrameDecompressor.decompress :135-210 */
int64_t orc_zstd_decompress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, int64_t *err_offset)
{
    if (out_cap == 0) return 0;                                                   /* :143-145 */
    static __thread zdctx *ctx;
    if (!ctx) { ctx = (zdctx *) __builtin_malloc(sizeof(zdctx)); }
    zdctx *c = ctx;
    zdctx_init(c);
    int64_t input = 0, output = 0;
    while (input < in_len) {
        c->prev[0] = 1; c->prev[1] = 4; c->prev[2] = 8;                            /* reset :212-221 */
        c->cur_ll = c->cur_of = c->cur_ml = 0;
        int64_t output_start = output;
        int64_t m = verify_magic(in, input, in_len, err_offset);
        if (m < 0) return m;
        input += m;
        frame_header fh;
        PROPAGATE(read_frame_header(&fh, in, input, in_len, err_offset));
        input += fh.header_size;
        int last_block;
        do {
            CHECK(input + 3 <= in_len, input, ZR_NOT_ENOUGH_INPUT);
            int32_t header = (int32_t) (zo_ld16(in + input) | ((uint32_t) in[input + 2] << 16));
            input += 3;
            last_block = header & 1;
            int block_type = (header >> 1) & 3;
            int32_t block_size = (header >> 3) & 0x1FFFFF;
            int64_t decoded;
            if (block_type == 0) {                                                  /* decodeRawBlock :223-229 */
                CHECK(input + block_size <= in_len, input, ZR_NOT_ENOUGH_INPUT);
                CHECK(output + block_size <= out_cap, input, ZR_OUTPUT_TOO_SMALL);
                memcpy(out + output, in + input, (size_t) block_size);
                decoded = block_size;
                input += block_size;
            }
            else if (block_type == 1) {                                             /* decodeRleBlock :231-263 */
                CHECK(input + 1 <= in_len, input, ZR_NOT_ENOUGH_INPUT);
                CHECK(output + block_size <= out_cap, input, ZR_OUTPUT_TOO_SMALL);
                memset(out + output, in[input], (size_t) block_size);
                decoded = block_size;
                input += 1;
            }
            else if (block_type == 2) {
                CHECK(input + block_size <= in_len, input, ZR_NOT_ENOUGH_INPUT);
                decoded = decode_compressed_block(c, in, input, block_size, out, output, out_cap, fh.window_size, err_offset);
                if (decoded < 0) return decoded;
                input += block_size;
            }
            else {
                FAIL(input, ZR_INVALID_BLOCK_TYPE);
            }
            output += decoded;
        }
        while (!last_block);
        if (fh.has_checksum) {                                                      /* :194-206 */
            uint64_t hash = orc_xxh64(out + output_start, output - output_start, 0);
            CHECK(input + 4 <= in_len, input, ZR_NOT_ENOUGH_INPUT);
            uint32_t checksum = zo_ld32(in + input);
            if (checksum != (uint32_t) hash) FAIL(input, ZR_BAD_CHECKSUM);
            input += 4;
        }
    }
    return output;
}
