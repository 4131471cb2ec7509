// zstd_common.cuh -- device-side Zstandard building blocks shared by the decode and encode kernels.
//
// Bit reader semantics follow the reference's BitInputStream (zstd/BitInputStream.java:34-205) exactly,
// because the Java decoder's accept/reject decisions ("Bit stream is not fully consumed", "Not all
// sequences were consumed") depend on its refill rules.  Table construction follows
// zstd/FseTableReader.java:111-159 and zstd/FseCompressionTable.java:138-154 (symbol spreading).
#pragma once
#include "acc_device.cuh"

namespace zs {

constexpr int kMaxBlock = 128 * 1024;

// reasons (status >> 8), texts in acc_api.cu / include/aircompress_cuda.h
enum : int {
    R_NOT_ENOUGH_INPUT = 32, R_INVALID_BLOCK_TYPE = 33, R_OUTPUT_TOO_SMALL = 34, R_CORRUPTED = 35, R_BAD_MAGIC = 36,
    R_V07_FORMAT = 37, R_DICTIONARY = 38, R_WINDOW_TOO_LARGE = 39, R_BLOCK_TOO_SMALL = 40, R_EXPECTED_TABLE = 41,
    R_DICTIONARY_CORRUPTED = 42, R_BLOCK_EXCEEDS_MAX = 44, R_OUTPUT_EXCEEDS_BLOCK = 45, R_VALUE_EXCEEDS_MAX = 46,
    R_BITSTREAM_EMPTY = 47, R_BITSTREAM_NO_END_MARK = 48, R_NOT_ALL_SEQUENCES = 49, R_BITSTREAM_NOT_CONSUMED = 50,
    R_FSE_TABLE_TOO_LARGE = 51, R_SYMBOL_TOO_LARGE = 52, R_TOO_MANY_SYMBOLS = 53, R_BAD_CHECKSUM = 54, R_FSE_OUTPUT_TOO_SMALL = 55
};

__device__ __forceinline__ int highbit(uint32_t v) { return v ? 31 - __clz(v) : -1; }

__device__ __forceinline__ uint64_t ld64u(const uint8_t *p) { return ld_u64_unaligned(p); }
__device__ __forceinline__ uint32_t ld32u(const uint8_t *p) { return ld_u32_unaligned(p); }
__device__ __forceinline__ uint32_t ld16u(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8); }

// ---- Constants.java:66-78, ZstdFrameDecompressor.java:68-83 --------------------------------------
static __device__ __constant__ uint8_t kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static __device__ __constant__ uint8_t kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                               1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static __device__ __constant__ int32_t kLLBase[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64,
                                               0x80, 0x100, 0x200, 0x400, 0x800, 0x1000, 0x2000, 0x4000, 0x8000, 0x10000};
static __device__ __constant__ int32_t kMLBase[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
                                               29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403,
                                               0x803, 0x1003, 0x2003, 0x4003, 0x8003, 0x10003};
// predefined distributions (SequenceEncoder.java:36-56)
static __device__ __constant__ int16_t kDefLL[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static __device__ __constant__ int16_t kDefML[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                              1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static __device__ __constant__ int16_t kDefOF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

__device__ __forceinline__ int32_t of_base(int code) { return code == 0 ? 0 : (code == 1 ? 1 : (int32_t) ((1u << code) - 3)); }

// ---- backward bit reader (one per thread) ---------------------------------------------------------
struct BitReader {
    const uint8_t *in;   // stream base (global)
    int64_t start, cur;
    uint64_t bits;
    int32_t consumed;
    int overflow;
};

__device__ __forceinline__ uint64_t peek_bits(int32_t consumed, uint64_t bits, int n) { return ((bits << (consumed & 63)) >> 1) >> ((63 - n) & 63); }
__device__ __forceinline__ uint64_t peek_bits_fast(int32_t consumed, uint64_t bits, int n) { return (bits << (consumed & 63)) >> ((64 - n) & 63); }

// returns 0 or a reason code
__device__ __forceinline__ int br_init(BitReader &b, const uint8_t *in, int64_t start, int64_t end, int64_t *err_off)
{
    if (end - start < 1) { *err_off = start; return R_BITSTREAM_EMPTY; }
    int last = in[end - 1];
    if (last == 0) { *err_off = end; return R_BITSTREAM_NO_END_MARK; }
    b.in = in; b.start = start; b.overflow = 0;
    b.consumed = 8 - highbit((uint32_t) last);
    int64_t size = end - start;
    if (size >= 8) { b.cur = end - 8; b.bits = ld64u(in + b.cur); }
    else {
        b.cur = start;
        uint64_t v = in[start];
        for (int i = 1; i < size; i++) v |= (uint64_t) in[start + i] << (8 * i);
        b.bits = v;
        b.consumed += (int32_t) (8 - size) * 8;
    }
    return 0;
}

__device__ __forceinline__ bool br_load(BitReader &b)
{
    if (b.consumed > 64) { b.overflow = 1; return true; }
    if (b.cur == b.start) return true;
    int32_t bytes = (int32_t) ((uint32_t) b.consumed >> 3);
    if (b.cur >= b.start + 8) {
        if (bytes > 0) { b.cur -= bytes; b.bits = ld64u(b.in + b.cur); }
        b.consumed &= 7;
    }
    else if (b.cur - bytes < b.start) {
        bytes = (int32_t) (b.cur - b.start);
        b.cur = b.start;
        b.consumed -= bytes * 8;
        b.bits = ld64u(b.in + b.start);
        return true;
    }
    else {
        b.cur -= bytes;
        b.consumed -= bytes * 8;
        b.bits = ld64u(b.in + b.cur);
    }
    return false;
}

// ---- FSE decode table entry: new_state (16) | nbits (8) << 16 | symbol (8) << 24 -------------------
__device__ __forceinline__ uint32_t fse_entry(int new_state, int nbits, int symbol)
{
    return ((uint32_t) new_state & 0xFFFF) | ((uint32_t) nbits << 16) | ((uint32_t) symbol << 24);
}

// Builds a decode table from normalized counters (single thread).  `scratch` holds table_size bytes.
// Returns false when the counters do not fill the table exactly ("Input is corrupted").
__device__ inline bool fse_build_dtable(uint32_t *table, const int16_t *norm, int max_symbol, int table_log, uint8_t *scratch, int16_t *next)
{
    const int size = 1 << table_log;
    int high = size - 1;
    for (int s = 0; s <= max_symbol; s++) {
        if (norm[s] == -1) { scratch[high--] = (uint8_t) s; next[s] = 1; }
        else next[s] = norm[s];
    }
    const int mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    int position = 0;
    for (int s = 0; s <= max_symbol; s++) {
        for (int i = 0; i < norm[s]; i++) {
            scratch[position] = (uint8_t) s;
            do { position = (position + step) & mask; } while (position > high);
        }
    }
    if (position != 0) return false;
    for (int i = 0; i < size; i++) {
        int s = scratch[i];
        int ns = next[s]++;
        int nb = table_log - highbit((uint32_t) ns);
        table[i] = fse_entry((ns << nb) - size, nb, s);
    }
    return true;
}

}  // namespace zs
