// zstd_dec.cu -- Zstandard frame decode kernel for sm_100a.
//
// Replaces ZstdFrameDecompressor.decompress (zstd/ZstdFrameDecompressor.java:135-210) and everything it
// calls: frame/block headers (:860-940, :163-190), literals (:708-858 + zstd/Huffman.java:52-324),
// sequence tables (:609-676 + zstd/FseTableReader.java:27-168), sequence decode + execution (:312-516),
// XXH64 frame checksum (:194-206).  Output and accept/reject decisions are bit-exact with the Java
// decoder (offsets in error reports are relative to the input start).
//
// Mapping: one warp per input (all frames of the input, all blocks of a frame, in order).  Per warp in
// shared memory: the Huffman decode table (4096 x u16), three FSE decode tables (512/512/256 x u32) and
// a 32-entry sequence batch.  Lane 0 parses headers and walks the FSE state chain, lanes 0-3 decode the
// four Huffman streams, all 32 lanes build tables and execute literal / match copies.
#include "zstd_common.cuh"
#include "xxh64_device.cuh"

namespace {
using namespace zs;

constexpr int kWarpsPerCta = 4;
constexpr int kSeqBatch = 32;
constexpr int kRing = 2048, kRingMask = kRing - 1;   // output ring of the sequence executor (bytes)
constexpr int kFlush = 512;                          // the ring drains in pieces of at least this many bytes
constexpr int32_t kFastMinBits = 256;                // the wide sequence path needs this many unread bits in front of a sequence
constexpr int kHufSmemLog = 11;                      // Huffman tables up to this log live in shared memory
constexpr int kSlots = 2;                            // record batches a worker holds (the chain lane of the service kernel runs ahead)

// Shared memory of a warp, 7.4 KiB in three regions whose tenants are never live at the same time:
//   A  the Huffman table while the literals of a block are decoded (2048 x u16: table logs up to 11, which is what the Java and
//      libzstd encoders emit; a 12-bit table -- legal for the Java decoder -- lives in the warp's global scratch), then the
//      three FSE tables while its sequences are decoded;
//   B  table-building scratch (weights, normalized counts, the FSE table of the Huffman weights), then the output ring;
//   C  the sequence batch of the wide path / of the exact loop.
// Tables a later block or frame of the same input may reuse (treeless literals, repeat-mode sequence tables) are parked in the
// warp's global scratch (kHufSave / kFseSave) whenever more input follows the current block.  The kernel is latency bound
// (serial bit-stream walks), so resident warps per SM are what this layout buys: 28 instead of 20 with 10.2 KiB.
struct WarpSmem {
    union {
        uint16_t huf[2048];                       // symbol | nbits << 8
        struct { uint32_t ll[512], ml[512], of[256]; };   // sequence tables after the fix-up pass: new_state (10) | 2 zero bits | extra bits of the code << 12 (7 wide) | nbits << 19 (5 wide) | symbol << 24: three entries ADD without a carry between fields
    };
    union {
        uint8_t ring[kRing];                      // the newest output bytes of the block being executed (16-byte aligned: offset 5120)
        struct {
            uint32_t wt[64];                      // FSE table of the Huffman weights
            int16_t norm[256];
            int16_t next[256];
            uint8_t scratch[512];                 // symbol spread buffer / Huffman weights
            int32_t ranks[16];
        };
    };
    union {
        uint2 rec[kSlots][kSeqBatch];             // wide path, per sequence: {unread bits in front of it, its three codes}; two batches
        struct { int32_t seq_ll[kSeqBatch], seq_ml[kSeqBatch], seq_of[kSeqBatch]; };   // exact loop
    };
};
static_assert(sizeof(WarpSmem) == 5120 + kRing + 512 && sizeof(WarpSmem) % 16 == 0, "WarpSmem layout");

#ifndef LZS_EMU
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#else
__device__ __forceinline__ void prefetch_l1(const void *) {}
#endif

// ---- the FSE state walk ------------------------------------------------------------------------------------------------
// One lane, one block: table entry -> bits consumed -> next state, noting where each sequence's bits begin.  Lengths and
// offsets are NOT assembled here (the worker's lanes do that in parallel); the next window of the stream is requested before
// it is needed.  Valid while at least kFastMinBits unread bits lie in front of a sequence.
struct Chain {
    const uint8_t *bs;        // first byte of the sequence bit stream
    int32_t P, wb;            // unread bits; byte position of the window
    uint64_t w;               // bytes [wb, wb + 8) of the stream
    uint32_t sl, sm, so;      // the three states
    __device__ __forceinline__ void open()
    {
        wb = (P - 57) >> 3;   // the window's top is 0..7 bits above P
        w = ld64u(bs + wb);
    }
    __device__ __forceinline__ uint2 step(const uint32_t *ll, const uint32_t *ml, const uint32_t *of)
    {
        const uint32_t el = ll[sl], em = ml[sm], eo = of[so];
        const uint32_t sum = el + em + eo;                // extra-bit total in bits 12-18, state-bit total in bits 19-23 (the fields cannot carry)
        const uint2 r = make_uint2((uint32_t) P, __byte_perm(__byte_perm(el, em, 0x4473), eo, 0x4710));   // the three codes (bytes 0-2)
        const int32_t P1 = P - (int32_t) ((sum >> 12) & 0x7F);   // behind the extra bits
        int32_t sft = P1 - wb * 8;                        // the state bits are bits [sft - 26, sft) of the window
        if (sft < 32) { wb = (P1 - 57) >> 3; w = ld64u(bs + wb); sft = P1 - wb * 8; }
        uint32_t x = __funnelshift_rc((uint32_t) w, (uint32_t) (w >> 32), (uint32_t) (sft - 32));     // bits [sft - 32, sft)
        const uint32_t nbl = (el >> 19) & 31, nbm = (em >> 19) & 31, nbo = (eo >> 19) & 31;
        sl = (el & 0x3FF) + __funnelshift_lc(x, 0, nbl); x <<= nbl;
        sm = (em & 0x3FF) + __funnelshift_lc(x, 0, nbm); x <<= nbm;
        so = (eo & 0x3FF) + __funnelshift_lc(x, 0, nbo);
        P = P1 - (int32_t) ((sum >> 19) & 0x1F);
        const int32_t nwb = (P - 57) >> 3;                // next window, requested now, needed one table lookup later
        if (((nwb ^ wb) & ~127) != 0 && nwb >= 384) prefetch_l1(bs + nwb - 384);   // the stream is read downwards: a new line every ~40 sequences
        wb = nwb;
        w = ld64u(bs + wb);
        return r;
    }
};

// ---- the state walk as a service ---------------------------------------------------------------------------------------
// The kernel is bound by warp instructions (ALU pipe), and the state walk is a third of them with ONE lane working.  In the
// service kernel a CTA is kWorkers worker warps + one chain warp whose lane w walks the states of worker w's block: one warp
// instruction advances up to kWorkers blocks.  Worker and lane talk through a mailbox in shared memory: the worker posts
// {stream, unread bits, states, sequence count}; the lane fills the worker's two record slots batch by batch (it runs ahead
// by at most two batches) and ends with the state it stopped in -- at the last sequence, or kFastMinBits before the start of
// the stream, where the worker's exact loop takes over.  A worker that gives up on a block (malformed input) simply posts its
// next request: batches carry the request number and stale ones are dropped.
struct ChainBox {
    const uint8_t *bs;
    int32_t P;
    int32_t n;                // sequences wanted
    uint32_t states;          // ll | ml << 10 | of << 20
    uint32_t posted;          // requests posted by the worker
    uint32_t consumed;        // batches the worker is done with
    uint32_t produced;        // batches published by the chain lane
    uint32_t count[kSlots];   // sequences of the batch | request number << 8 | kChainFinal
    int32_t end_P;            // with the final batch: where the walk stopped
    uint32_t end_states;
};
constexpr uint32_t kChainFinal = 0x80000000u;

#ifndef LZS_EMU
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
__device__ __forceinline__ void pause_ns(unsigned ns) { __nanosleep(ns); }
#else
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
__device__ __forceinline__ void pause_ns(unsigned) { sched_yield(); }
#endif

// the chain warp of a CTA: lane w serves worker w in rounds of one sequence per active lane
template <int kWorkers>
__device__ void chain_warp(ChainBox *boxes, const struct WarpSmem *sms, const uint32_t *workers_done, const int lane)
{
    ChainBox *const box = boxes + (lane < kWorkers ? lane : 0);
    const WarpSmem &sm = sms[lane < kWorkers ? lane : 0];
    uint2 *const slots = const_cast<uint2 *>(&sm.rec[0][0]);
    const bool serving = lane < kWorkers;
    bool active = false;
    uint32_t seen = 0, batches = 0, fill = 0;
    int32_t n = 0;
    Chain ch;
    ch.bs = nullptr; ch.P = 0; ch.wb = 0; ch.w = 0; ch.sl = ch.sm = ch.so = 0;
    for (;;) {
        bool worked = false;
        if (serving) {
            const uint32_t posted = ld_vol(&box->posted);
            if (posted != seen) {                          // a new request (also: the worker gave up on the previous one)
                seen = posted;
                __threadfence_block();
                ch.bs = box->bs; ch.P = box->P; n = box->n;
                const uint32_t st = box->states;
                ch.sl = st & 1023; ch.sm = (st >> 10) & 1023; ch.so = st >> 20;
                if (ch.P >= kFastMinBits) ch.open();
                active = true;
                fill = 0;
            }
            bool go = active;
            if (go && fill == 0) go = batches - ld_vol(&box->consumed) < (uint32_t) kSlots;   // a free slot?
            if (go) {
                uint2 *const slot = slots + (batches & (kSlots - 1)) * kSeqBatch;
                if (ch.P >= kFastMinBits && n > 0) { slot[fill++] = ch.step(sm.ll, sm.ml, sm.of); n--; }
                const bool fin = n == 0 || ch.P < kFastMinBits;
                if (fill == (uint32_t) kSeqBatch || fin) {
                    if (fin) { box->end_P = ch.P; box->end_states = ch.sl | (ch.sm << 10) | (ch.so << 20); active = false; }
                    box->count[batches & (kSlots - 1)] = fill | ((seen & 0x7FFFFFu) << 8) | (fin ? kChainFinal : 0u);
                    __threadfence_block();
                    st_vol(&box->produced, ++batches);
                    fill = 0;
                }
                worked = true;
            }
        }
        if (!__any_sync(kFull, worked)) {
            if (__all_sync(kFull, ld_vol(workers_done) == (uint32_t) kWorkers)) return;   // a uniform decision: all lanes leave together
            pause_ns(200);
        }
    }
}


struct Ctl {   // lane-0 results broadcast through registers
    int32_t reason;
    int64_t err_off;
};

#define ZFAIL(reason_, off_) do { ctl.reason = (reason_); ctl.err_off = (off_); return -1; } while (0)
#define ZCHECK(cond, off_, reason_) do { if (!(cond)) ZFAIL(reason_, off_); } while (0)

// FseTableReader.readFseTable :27-160 (single thread).  Returns bytes consumed or -1.
__device__ int64_t fse_read_table(uint32_t *table, int *table_log_out, const uint8_t *in, int64_t in_addr, int64_t in_limit, int max_symbol,
                                  int max_table_log, WarpSmem &sm, Ctl &ctl)
{
    int16_t *norm = sm.norm;
    int64_t input = in_addr;
    ZCHECK(in_limit - in_addr >= 4, input, R_NOT_ENOUGH_INPUT);
    int symbol_number = 0;
    bool previous_is_zero = false;
    uint32_t bit_stream = ld32u(in + input);
    int table_log = (int) (bit_stream & 0xF) + 5;
    int nbits = table_log + 1;
    bit_stream >>= 4;
    int bit_count = 4;
    ZCHECK(table_log <= max_table_log, input, R_FSE_TABLE_TOO_LARGE);
    int remaining = (1 << table_log) + 1;
    int threshold = 1 << table_log;
    while (remaining > 1 && symbol_number <= max_symbol) {
        if (previous_is_zero) {
            int n0 = symbol_number;
            while ((bit_stream & 0xFFFF) == 0xFFFF) {
                n0 += 24;
                if (input < in_limit - 5) { input += 2; bit_stream = ld32u(in + input) >> (bit_count & 31); }
                else { bit_stream >>= 16; bit_count += 16; }
            }
            while ((bit_stream & 3) == 3) { n0 += 3; bit_stream >>= 2; bit_count += 2; }
            n0 += (int) (bit_stream & 3);
            bit_count += 2;
            ZCHECK(n0 <= max_symbol, input, R_SYMBOL_TOO_LARGE);
            while (symbol_number < n0) norm[symbol_number++] = 0;
            if (input <= in_limit - 7 || input + (bit_count >> 3) <= in_limit - 4) {
                input += bit_count >> 3;
                bit_count &= 7;
                bit_stream = ld32u(in + input) >> bit_count;
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
        bit_stream = ld32u(in + input) >> (bit_count & 31);
    }
    ZCHECK(remaining == 1 && bit_count <= 32, input, R_CORRUPTED);
    int max_sym = symbol_number - 1;
    ZCHECK(max_sym <= 255, input, R_TOO_MANY_SYMBOLS);
    input += (bit_count + 7) >> 3;
    if (!fse_build_dtable(table, norm, max_sym, table_log, sm.scratch, sm.next)) ZFAIL(R_CORRUPTED, input);
    *table_log_out = table_log;
    return input - in_addr;
}

// FiniteStateEntropy.decompress :38-151 (Huffman weights, single thread).  Returns symbol count or -1.
__device__ int fse_decompress_weights(const uint32_t *table, int log2, const uint8_t *in, int64_t in_addr, int64_t in_limit, uint8_t *out, int out_cap,
                                      Ctl &ctl)
{
    BitReader b;
    int64_t eo = 0;
    int r = br_init(b, in, in_addr, in_limit, &eo);
    if (r) ZFAIL(r, eo);
    int output = 0;
    int state1 = (int) peek_bits(b.consumed, b.bits, log2); b.consumed += log2;
    br_load(b);
    int state2 = (int) peek_bits(b.consumed, b.bits, log2); b.consumed += log2;
    br_load(b);
#define FSE_SYM(st) (uint8_t) (table[st] >> 24)
#define FSE_STEP(st) do { uint32_t e_ = table[st]; int nb_ = (e_ >> 16) & 0xFF; st = (int) ((e_ & 0xFFFF) + (uint32_t) peek_bits(b.consumed, b.bits, nb_)); b.consumed += nb_; } while (0)
    while (output <= out_cap - 4) {
        out[output] = FSE_SYM(state1); FSE_STEP(state1);
        out[output + 1] = FSE_SYM(state2); FSE_STEP(state2);
        out[output + 2] = FSE_SYM(state1); FSE_STEP(state1);
        out[output + 3] = FSE_SYM(state2); FSE_STEP(state2);
        output += 4;
        if (br_load(b)) break;
    }
    for (;;) {
        ZCHECK(output <= out_cap - 2, in_addr, R_FSE_OUTPUT_TOO_SMALL);
        out[output++] = FSE_SYM(state1); FSE_STEP(state1);
        br_load(b);
        if (b.overflow) { out[output++] = FSE_SYM(state2); break; }
        ZCHECK(output <= out_cap - 2, in_addr, R_FSE_OUTPUT_TOO_SMALL);
        out[output++] = FSE_SYM(state2); FSE_STEP(state2);
        br_load(b);
        if (b.overflow) { out[output++] = FSE_SYM(state1); break; }
    }
#undef FSE_STEP
#undef FSE_SYM
    return output;
}

struct FrameState {
    int huf_log;              // -1 = no Huffman table loaded
    int ll_log, of_log, ml_log;   // -1 = no table yet in this frame
    int32_t prev[3];
};

// Huffman.readTable :52-128.  Lane 0 reads the weights, all lanes fill the table.  Returns bytes consumed or -1 (uniform).
__device__ int64_t huf_read_table(WarpSmem &sm, FrameState &fs, const uint8_t *in, int64_t in_addr, int size, uint16_t *big_table, Ctl &ctl, int lane)
{
    uint8_t *weights = sm.scratch;   // 257 needed; scratch has 512
    int32_t *ranks = sm.ranks;
    int64_t ret = 0;
    int number_of_symbols = 0, table_log = 0;
    if (lane == 0) {
        ret = [&]() -> int64_t {
            for (int i = 0; i < 16; i++) ranks[i] = 0;
            int64_t input = in_addr;
            ZCHECK(size > 0, input, R_NOT_ENOUGH_INPUT);
            int input_size = in[input++];
            int output_size;
            if (input_size >= 128) {
                output_size = input_size - 127;
                input_size = (output_size + 1) / 2;
                ZCHECK(input_size + 1 <= size, input, R_NOT_ENOUGH_INPUT);
                ZCHECK(output_size <= 256, input, R_CORRUPTED);
                for (int i = 0; i < output_size; i += 2) {
                    int v = in[input + i / 2];
                    weights[i] = (uint8_t) (v >> 4);
                    weights[i + 1] = (uint8_t) (v & 15);
                }
            }
            else {
                ZCHECK(input_size + 1 <= size, input, R_NOT_ENOUGH_INPUT);
                int64_t limit = input + input_size;
                int wlog = 0;
                int64_t used = fse_read_table(sm.wt, &wlog, in, input, limit, 255, 6, sm, ctl);
                if (used < 0) return -1;
                input += used;
                int n = fse_decompress_weights(sm.wt, wlog, in, input, limit, weights, 256, ctl);
                if (n < 0) return -1;
                output_size = n;
                ZCHECK(output_size <= 255, input, R_CORRUPTED);   // the Java indexes weights[outputSize] in a byte[256]
            }
            int total_weight = 0;
            for (int i = 0; i < output_size; i++) {
                ZCHECK(weights[i] <= 12, input, R_CORRUPTED);
                ranks[weights[i]]++;
                total_weight += (1 << weights[i]) >> 1;
            }
            ZCHECK(total_weight != 0, input, R_CORRUPTED);
            int tl = highbit((uint32_t) total_weight) + 1;
            ZCHECK(tl <= 12, input, R_CORRUPTED);
            int rest = (1 << tl) - total_weight;
            ZCHECK((rest & (rest - 1)) == 0, input, R_CORRUPTED);
            int last_weight = highbit((uint32_t) rest) + 1;
            weights[output_size] = (uint8_t) last_weight;
            ranks[last_weight]++;
            int next_rank_start = 0;
            for (int i = 1; i < tl + 1; ++i) {
                int current = next_rank_start;
                next_rank_start += ranks[i] << (i - 1);
                ranks[i] = current;
            }
            // per-symbol start positions (serial prefix in symbol order, like the Java loop); stored in norm[]
            int r1_end = 0;
            for (int n = 0; n < output_size + 1; n++) {
                int w = weights[n];
                sm.norm[n] = (int16_t) ranks[w];
                ranks[w] += (1 << w) >> 1;
            }
            r1_end = ranks[1];
            ZCHECK(r1_end >= 2 && (r1_end & 1) == 0, input, R_CORRUPTED);
            number_of_symbols = output_size + 1;
            table_log = tl;
            return input_size + 1;
        }();
    }
    ret = __shfl_sync(kFull, ret, 0);
    if (ret < 0) { ctl.reason = __shfl_sync(kFull, ctl.reason, 0); ctl.err_off = __shfl_sync(kFull, ctl.err_off, 0); return -1; }
    number_of_symbols = __shfl_sync(kFull, number_of_symbols, 0);
    table_log = __shfl_sync(kFull, table_log, 0);
    __syncwarp();
    // fill: symbol n covers [start, start + (1 << w) >> 1)
    uint16_t *const tab = table_log > kHufSmemLog ? big_table : sm.huf;
    for (int n = 0; n < number_of_symbols; n++) {
        int w = weights[n];
        int length = (1 << w) >> 1;
        int start = (uint16_t) sm.norm[n];
        uint16_t e = (uint16_t) (n | ((table_log + 1 - w) << 8));
        for (int i = lane; i < length; i += 32) tab[start + i] = e;
    }
    __syncwarp();
    fs.huf_log = table_log;
    return ret;
}

// One Huffman stream decoded by one thread (Huffman.decodeTail semantics, zstd/Huffman.java:291-317).  Away from the start
// of the stream a refill is BitInputStream.load() on its common path (at least 57 unread bits afterwards), which four symbols
// of at most 12 bits cannot exhaust: the main loop decodes four symbols per refill from a top-aligned copy of the word and
// writes them with one 32-bit store.  The last bytes of the stream and of the output go through the symbol-by-symbol loops.
__device__ int huf_decode_stream(const uint16_t *huf, int tl, const uint8_t *in, int64_t start, int64_t end, uint8_t *out, int64_t n, int64_t *err_off)
{
    BitReader b;
    int r = br_init(b, in, start, end, err_off);
    if (r) return r;
    int64_t o = 0;
    bool at_start = false;
    while (o < n && ((uintptr_t) (out + o) & 3)) {           // up to the first 4-byte aligned output address
        if (br_load(b)) { at_start = true; break; }
        uint32_t e = huf[(int) peek_bits_fast(b.consumed, b.bits, tl)];
        out[o++] = (uint8_t) e;
        b.consumed += e >> 8;
    }
    if (!at_start) {
        const int sh = 64 - tl;
        while (o + 4 <= n && b.cur >= b.start + 8) {
            b.cur -= b.consumed >> 3;
            b.consumed &= 7;
            b.bits = ld64u(b.in + b.cur);
            uint64_t v = b.bits << b.consumed;
            const uint32_t e0 = huf[(uint32_t) (v >> sh)]; v <<= e0 >> 8;
            const uint32_t e1 = huf[(uint32_t) (v >> sh)]; v <<= e1 >> 8;
            const uint32_t e2 = huf[(uint32_t) (v >> sh)]; v <<= e2 >> 8;
            const uint32_t e3 = huf[(uint32_t) (v >> sh)];
            b.consumed += (int32_t) ((e0 >> 8) + (e1 >> 8) + (e2 >> 8) + (e3 >> 8));
            *reinterpret_cast<uint32_t *>(out + o) = (e0 & 0xFF) | ((e1 & 0xFF) << 8) | ((e2 & 0xFF) << 16) | (e3 << 24);
            o += 4;
        }
        while (o < n) {
            if (br_load(b)) break;
            uint32_t e = huf[(int) peek_bits_fast(b.consumed, b.bits, tl)];
            out[o++] = (uint8_t) e;
            b.consumed += e >> 8;
        }
    }
    while (o < n) {
        uint32_t e = huf[(int) peek_bits_fast(b.consumed, b.bits, tl)];
        out[o++] = (uint8_t) e;
        b.consumed += e >> 8;
    }
    if (!(b.start == b.cur && b.consumed == 64)) { *err_off = start; return R_BITSTREAM_NOT_CONSUMED; }
    return 0;
}

struct Literals {
    const uint8_t *ptr;   // raw / decoded literals; unused for RLE
    int64_t size;
    int rle;              // -1 = not RLE, else the byte
};

// copy `n` literal bytes starting at literal position `pos` to dst (warp-cooperative)
__device__ __forceinline__ void copy_literals(uint8_t *dst, const Literals &lit, int64_t pos, int64_t n, int lane)
{
    if (lit.rle >= 0) {
        for (int64_t i = lane; i < n; i += 32) dst[i] = (uint8_t) lit.rle;
    }
    else {
        warp_copy(dst, lit.ptr + pos, n, lane);
    }
}

// A worker that leaves an input early (malformed) must not reuse its tables while its chain lane may still be walking them:
// it posts an empty request and waits for that request's (empty, final) batch -- the lane is idle from then on.  (Called from
// the kernel loop, not from the decode functions: a call in there cost the hot loops registers, -20 %.)
__device__ __noinline__ void chain_cancel(ChainBox *box, const int lane)
{
    uint32_t my_batches = box->consumed;
    __syncwarp();
    if (lane == 0) {
        box->P = 0; box->n = 0;
        __threadfence_block();
        st_vol(&box->posted, box->posted + 1);
    }
    __syncwarp();
    const uint32_t my_req = box->posted;
    for (;;) {
        while ((int32_t) (ld_vol(&box->produced) - my_batches) <= 0) pause_ns(100);
        __threadfence_block();
        const uint32_t c = box->count[my_batches & (kSlots - 1)];
        __syncwarp();
        my_batches++;
        if (lane == 0) st_vol(&box->consumed, my_batches);
        if (((c >> 8) & 0x7FFFFFu) == (my_req & 0x7FFFFFu)) break;
    }
    __syncwarp();
}

// ---- the wide sequence path: output ring -------------------------------------------------------------------------------
// Positions are 32-bit and relative to the block: output position inside the block + (address of its first output byte & 15),
// so that ring units and global 16-byte units are aligned alike; `out_al` is the (16-byte aligned) address of position 0.  A
// match source in front of the block has a negative position.  `ring_lo` is the first position the ring holds.

// moves [flushed, e) from the ring to global memory: 16-byte stores for whole aligned units, bytes for the partial units at both ends
__device__ __forceinline__ void ring_flush(const uint8_t *ring, uint8_t *out_al, int32_t &flushed, const int32_t e, const int lane)
{
    int32_t a = flushed;
    if (a & 15) {
        int32_t a1 = (a + 15) & ~15;
        if (a1 > e) a1 = e;
        if (a + lane < a1) out_al[a + lane] = ring[(a + lane) & kRingMask];
        a = a1;
    }
    const int32_t units = (e - a) >> 4;
    for (int32_t u = lane; u < units; u += 32) {
        const int32_t w = a + (u << 4);
        *reinterpret_cast<uint4 *>(out_al + w) = *reinterpret_cast<const uint4 *>(ring + (w & kRingMask));
    }
    a += units << 4;
    if (a + lane < e) out_al[a + lane] = ring[(a + lane) & kRingMask];
    flushed = e;
    __syncwarp();
}

// one sequence the multi-sequence step cannot take (more than 64 bytes, or a match that overlaps its own output): its literals,
// then the match, in pieces of at most kFlush bytes so that the ring can drain in between
__device__ __forceinline__ void ring_long_sequence(uint8_t *ring, const uint8_t *lit, uint8_t *out_al, int32_t &flushed, const int32_t ring_lo,
                                                   const int32_t opw, const int32_t ll, const int32_t ml, const int32_t off, const int lane)
{
    for (int32_t cb = 0; cb < ll; cb += kFlush) {
        const int32_t pn = ll - cb < kFlush ? ll - cb : kFlush;
        for (int32_t i = lane; i < pn; i += 32) ring[(opw + cb + i) & kRingMask] = lit[cb + i];
        __syncwarp();
        if (opw + cb + pn - flushed >= kFlush) ring_flush(ring, out_al, flushed, (opw + cb + pn) & ~15, lane);
    }
    const int32_t mopw = opw + ll;
    for (int32_t cb = 0; cb < ml; cb += kFlush) {
        const int32_t pw = mopw + cb;                                   // first byte of this piece
        const int32_t pn = ml - cb < kFlush ? ml - cb : kFlush;
        if (off >= 32) {
            for (int32_t base = 0; base < pn; base += 32) {             // a round only reads bytes written at least 32 positions earlier
                const int32_t i = base + lane;
                if (i < pn) {
                    const int32_t pos = pw + i, src = pos - off;
                    const int32_t lo = pw + base + 32 - kRing > ring_lo ? pw + base + 32 - kRing : ring_lo;
                    ring[pos & kRingMask] = src >= lo ? ring[src & kRingMask] : out_al[src];   // older than the ring: flushed long ago
                }
                __syncwarp();
            }
        }
        else {
            // periodic pattern: every byte of the piece repeats one of the `off` bytes in front of it
            int32_t m = lane % off;
            const int32_t step = 32 % off;
            for (int32_t i = lane; i < pn; i += 32) {
                const int32_t src = pw - off + m;
                ring[(pw + i) & kRingMask] = src >= ring_lo ? ring[src & kRingMask] : out_al[src];
                m += step;
                if (m >= off) m -= off;
            }
            __syncwarp();
        }
        if (pw + pn - flushed >= kFlush) ring_flush(ring, out_al, flushed, (pw + pn) & ~15, lane);
    }
}

// decodes one compressed block (ZstdFrameDecompressor.decodeCompressedBlock :265-310 + decompressSequences :312-516).
// Returns bytes produced or -1.  `out`/`out_pos` are relative to the start of the caller's output buffer.
// copies n_bytes (a multiple of 16) between 16-byte aligned buffers with the whole warp
__device__ __forceinline__ void warp_copy16(void *dst, const void *src, int n_bytes, int lane)
{
    uint4 *d = (uint4 *) dst;
    const uint4 *q = (const uint4 *) src;
    for (int i = lane; i < n_bytes / 16; i += 32) d[i] = q[i];
    __syncwarp();
}

constexpr int kHufBytes = 4096 * 2, kFseBytes = (512 + 512 + 256) * 4;
constexpr int64_t kHufSave = kMaxBlock + 256, kFseSave = kHufSave + kHufBytes;   // offsets in the warp's scratch

template <bool kSvc>
__device__ int64_t decode_compressed_block(WarpSmem &sm, ChainBox *box, FrameState &fs, const uint8_t *in, int64_t in_addr, int block_size, uint8_t *out,
                                           int64_t out_pos, int64_t out_cap, int32_t window_size, uint8_t *lit_scratch, bool keep_tables,
                                           Ctl &ctl, int lane)
{
    int64_t input = in_addr;
    const int64_t block_end = in_addr + block_size;
    ZCHECK(block_size <= kMaxBlock, input, R_EXPECTED_TABLE);
    ZCHECK(block_size >= 3, input, R_BLOCK_TOO_SMALL);
    Literals lit;
    lit.rle = -1; lit.ptr = nullptr; lit.size = 0;
    const int b0 = in[input];
    const int lit_type = b0 & 3;
    const int size_format = (b0 >> 2) & 3;
    if (lit_type == 0 || lit_type == 1) {
        // decodeRawLiterals :812-858 / decodeRleLiterals :776-810
        int32_t lsize;
        if (size_format == 0 || size_format == 2) { lsize = b0 >> 3; input += 1; }
        else if (size_format == 1) { lsize = (int32_t) (ld16u(in + input) >> 4); input += 2; }
        else {
            if (lit_type == 1) ZCHECK(block_size >= 4, input, R_NOT_ENOUGH_INPUT);
            lsize = (int32_t) (((uint32_t) in[input] | (ld16u(in + input + 1) << 8)) >> 4);
            input += 3;
        }
        if (lit_type == 0) {
            ZCHECK(input + lsize <= block_end, input, R_NOT_ENOUGH_INPUT);
            lit.ptr = in + input;
            lit.size = lsize;
            input += lsize;
        }
        else {
            ZCHECK(lsize <= kMaxBlock, input, R_OUTPUT_EXCEEDS_BLOCK);
            lit.rle = in[input++];
            lit.size = lsize;
        }
    }
    else {
        if (lit_type == 3) ZCHECK(fs.huf_log != -1, input, R_DICTIONARY_CORRUPTED);
        // decodeCompressedLiterals :708-774
        ZCHECK(block_size >= 5, input, R_NOT_ENOUGH_INPUT);
        int32_t comp_size, unc_size, header_size;
        bool single = false;
        if (size_format == 0 || size_format == 1) {
            single = size_format == 0;
            uint32_t hd = ld32u(in + input);
            header_size = 3; unc_size = (int32_t) ((hd >> 4) & 0x3FF); comp_size = (int32_t) ((hd >> 14) & 0x3FF);
        }
        else if (size_format == 2) {
            uint32_t hd = ld32u(in + input);
            header_size = 4; unc_size = (int32_t) ((hd >> 4) & 0x3FFF); comp_size = (int32_t) ((hd >> 18) & 0x3FFF);
        }
        else {
            uint64_t hd = (uint64_t) in[input] | ((uint64_t) ld32u(in + input + 1) << 8);
            header_size = 5; unc_size = (int32_t) ((hd >> 4) & 0x3FFFF); comp_size = (int32_t) ((hd >> 22) & 0x3FFFF);
        }
        ZCHECK(unc_size <= kMaxBlock, input, R_BLOCK_EXCEEDS_MAX);
        ZCHECK(header_size + comp_size <= block_size, input, R_CORRUPTED);
        input += header_size;
        const int64_t lit_limit = input + comp_size;
        uint16_t *const big_table = reinterpret_cast<uint16_t *>(lit_scratch + kHufSave);   // where a 12-bit table lives, and where smaller ones are parked
        if (lit_type != 3) {
            int64_t used = huf_read_table(sm, fs, in, input, comp_size, big_table, ctl, lane);
            if (used < 0) return -1;
            input += used;
            if (keep_tables && fs.huf_log <= kHufSmemLog) warp_copy16(big_table, sm.huf, (int) sizeof(sm.huf), lane);
        }
        else if (fs.huf_log <= kHufSmemLog) {
            // treeless literals: the table of an earlier block (parked when it was built: something followed that block)
            __syncwarp();
            warp_copy16(sm.huf, big_table, (int) sizeof(sm.huf), lane);
        }
        const uint16_t *const huf = fs.huf_log > kHufSmemLog ? big_table : sm.huf;
        // streams
        int reason = 0;
        int64_t eo = 0;
        const int tl = fs.huf_log;
        if (single) {
            if (lane == 0) reason = huf_decode_stream(huf, tl, in, input, lit_limit, lit_scratch, unc_size, &eo);
        }
        else {
            // decode4Streams :166-289
            if (lit_limit - input < 10) { reason = R_CORRUPTED; eo = input; }
            else {
                int64_t s1 = input + 6;
                int64_t s2 = s1 + ld16u(in + input), s3 = s2 + ld16u(in + input + 2), s4 = s3 + ld16u(in + input + 4);
                if (!(s2 < s3 && s3 < s4 && s4 < lit_limit)) { reason = R_CORRUPTED; eo = input; }
                else if (lane < 4) {
                    int64_t seg = ((int64_t) unc_size + 3) / 4;
                    int64_t st = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : s4;
                    int64_t en = lane == 0 ? s2 : lane == 1 ? s3 : lane == 2 ? s4 : lit_limit;
                    int64_t o0 = seg * lane;
                    int64_t n = lane < 3 ? seg : (int64_t) unc_size - 3 * seg;
                    if (n < 0) n = 0;   // tiny 4-stream sections: the Java decodes nothing from stream 4 but still wants it fully consumed
                    reason = huf_decode_stream(huf, tl, in, st, en, lit_scratch + o0, n, &eo);
                }
            }
        }
        // first failing lane (lowest stream index) wins, like the Java's stream-by-stream tail checks
        unsigned bad = __ballot_sync(kFull, reason != 0);
        if (bad) {
            int src = __ffs(bad) - 1;
            ctl.reason = __shfl_sync(kFull, reason, src);
            ctl.err_off = __shfl_sync(kFull, eo, src);
            return -1;
        }
        __syncwarp();
        lit.ptr = lit_scratch;
        lit.size = unc_size;
        input = lit_limit;
    }
    ZCHECK(window_size <= (1 << 23), input, R_WINDOW_TOO_LARGE);

    // ---- sequences section ----
    int64_t output = out_pos;
    int64_t lit_pos = 0;
    ZCHECK(block_end - input >= 1, input, R_NOT_ENOUGH_INPUT);
    int32_t seq_count = in[input++];
    if (seq_count != 0) {
        if (seq_count == 255) {
            ZCHECK(input + 2 <= block_end, input, R_NOT_ENOUGH_INPUT);
            seq_count = (int32_t) ld16u(in + input) + 0x7F00;
            input += 2;
        }
        else if (seq_count > 127) {
            ZCHECK(input < block_end, input, R_NOT_ENOUGH_INPUT);
            seq_count = ((seq_count - 128) << 8) + in[input++];
        }
        ZCHECK(input + 4 <= block_end, input, R_NOT_ENOUGH_INPUT);
        const uint32_t type = in[input++];
        // computeLiteralsTable / computeOffsetsTable / computeMatchLengthTable :609-676 (lane 0 builds, result broadcast)
        int64_t tb_ret = 0;
        int lg[3] = {fs.ll_log, fs.of_log, fs.ml_log};
        __syncwarp();   // the literal streams are done with the Huffman table: the region now holds the FSE tables
        if ((type >> 6) == 3 || ((type >> 4) & 3) == 3 || ((type >> 2) & 3) == 3) {
            // repeat mode: bring back the tables of the previous sequence section of this frame (if there is none the
            // lg[k] >= 0 check below rejects the block before anything is used)
            warp_copy16(sm.ll, lit_scratch + kFseSave, kFseBytes, lane);
        }
        if (lane == 0) {
            tb_ret = [&]() -> int64_t {
                const int types[3] = {(int) (type >> 6), (int) ((type >> 4) & 3), (int) ((type >> 2) & 3)};
                uint32_t *tables[3] = {sm.ll, sm.of, sm.ml};
                const int max_sym[3] = {35, 28, 52};
                const int max_log[3] = {9, 8, 9};
                const int def_log[3] = {6, 5, 6};
                for (int k = 0; k < 3; k++) {
                    if (types[k] == 1) {
                        ZCHECK(input < block_end, input, R_NOT_ENOUGH_INPUT);
                        int8_t value = (int8_t) in[input++];
                        ZCHECK(value <= max_sym[k] && value >= 0, input, R_VALUE_EXCEEDS_MAX);
                        tables[k][0] = fse_entry(0, 0, value);
                        lg[k] = 0;
                    }
                    else if (types[k] == 0) {
                        const int16_t *def = k == 0 ? kDefLL : k == 1 ? kDefOF : kDefML;
                        for (int s = 0; s <= max_sym[k]; s++) sm.norm[s] = def[s];
                        fse_build_dtable(tables[k], sm.norm, max_sym[k], def_log[k], sm.scratch, sm.next);
                        lg[k] = def_log[k];
                    }
                    else if (types[k] == 3) {
                        ZCHECK(lg[k] >= 0, input, R_EXPECTED_TABLE);
                    }
                    else {
                        int tlog = 0;
                        int64_t used = fse_read_table(tables[k], &tlog, in, input, block_end, max_sym[k], max_log[k], sm, ctl);
                        if (used < 0) return -1;
                        input += used;
                        lg[k] = tlog;
                    }
                }
                return input;
            }();
        }
        tb_ret = __shfl_sync(kFull, tb_ret, 0);
        if (tb_ret < 0) { ctl.reason = __shfl_sync(kFull, ctl.reason, 0); ctl.err_off = __shfl_sync(kFull, ctl.err_off, 0); return -1; }
        input = tb_ret;
        fs.ll_log = __shfl_sync(kFull, lg[0], 0);
        fs.of_log = __shfl_sync(kFull, lg[1], 0);
        fs.ml_log = __shfl_sync(kFull, lg[2], 0);
        __syncwarp();
        // re-pack the entries (layout at WarpSmem): the number of extra bits of every code goes in, so the state walk needs no
        // second lookup, and the fields are placed so that the three entries of a sequence add without carries.  Tables brought back from the parking area already carry them.
        for (int k = 0; k < 3; k++) {
            if (((type >> (6 - 2 * k)) & 3) == 3) continue;
            uint32_t *tab = k == 0 ? sm.ll : k == 1 ? sm.of : sm.ml;
            const int size = 1 << (k == 0 ? fs.ll_log : k == 1 ? fs.of_log : fs.ml_log);
            for (int i = lane; i < size; i += 32) {
                const uint32_t e = tab[i], sym = e >> 24;
                const uint32_t ext = k == 0 ? (uint32_t) kLLBits[sym] : k == 1 ? sym : (uint32_t) kMLBits[sym];
                tab[i] = (e & 0x3FF) | (ext << 12) | (((e >> 16) & 0x1F) << 19) | (e & 0xFF000000u);
            }
        }
        __syncwarp();
        if (keep_tables) warp_copy16(lit_scratch + kFseSave, sm.ll, kFseBytes, lane);

        // lane 0 owns the bit reader and the three FSE states; sequences are produced in batches of 32 and
        // executed by the whole warp.
        BitReader b;
        int ll_state = 0, of_state = 0, ml_state = 0;
        int32_t p0 = fs.prev[0], p1 = fs.prev[1], p2 = fs.prev[2];
        int init_reason = 0;
        int64_t init_eo = 0;
        if (lane == 0) {
            init_reason = br_init(b, in, input, block_end, &init_eo);
            if (!init_reason) {
                ll_state = (int) peek_bits(b.consumed, b.bits, fs.ll_log); b.consumed += fs.ll_log;
                of_state = (int) peek_bits(b.consumed, b.bits, fs.of_log); b.consumed += fs.of_log;
                ml_state = (int) peek_bits(b.consumed, b.bits, fs.ml_log); b.consumed += fs.ml_log;
            }
        }
        init_reason = __shfl_sync(kFull, init_reason, 0);
        if (init_reason) { ctl.reason = init_reason; ctl.err_off = __shfl_sync(kFull, init_eo, 0); return -1; }

        int32_t remaining = seq_count;
        bool stop = false;   // overflow with sequenceCount == 0 -> leave the loop (Java `break`)

        // ---- the wide path (every block whose literals are not a single repeated byte, as long as the bit stream is not
        // about to run out).  Per batch of 32 sequences:
        //   lane 0     walks the three FSE states only: table entry -> bits consumed -> next state, and notes where each
        //              sequence's bits begin (no length or offset is assembled on the serial chain; the next window of the
        //              stream is requested before it is needed);
        //   all lanes  pull the extra bits of THEIR sequence out of the stream, resolve the repeated-offset codes in order,
        //              find output / literal positions with two prefix sums, check the three conditions the Java checks per
        //              sequence, and execute in STEPS of up to 64 output bytes covering as many consecutive sequences as
        //              end inside them and copy from in front of the step.  Output is built in a 2 KiB shared-memory ring
        //              (what a warp has just written is not in L1) and leaves it in 16-byte stores.
        // While more than kFastMinBits unread bits lie in front of a sequence no refill rule of BitInputStream can trigger, so
        // plain "bits in order" is what the Java reads; nearer to the start of the stream the exact loop below takes over.
        bool wide = lit.rle < 0;
        const int64_t abs0 = output;                                     // output position of ring position `oh`
        const int32_t oh = (int32_t) ((uintptr_t) (out + abs0) & 15);
        uint8_t *const out_al = out + abs0 - oh;
        const int32_t ring_lo = oh;
        int32_t opw = oh, flushed = oh;
        const uint8_t *const bs = in + input;                            // first byte of the sequence bit stream
        Chain ch;                                                        // lane 0 (inline walk)
        ch.bs = bs; ch.P = 0; ch.wb = 0; ch.w = 0;
        ch.sl = (uint32_t) ll_state; ch.sm = (uint32_t) ml_state; ch.so = (uint32_t) of_state;
        bool b_stale = false;                                            // `b` and the three states are behind the walk
        if (lane == 0) ch.P = (int32_t) (b.cur - b.start) * 8 + 64 - b.consumed;
        const uint32_t le_mask = 0xffffffffu >> (31 - lane);             // lanes <= mine
        uint32_t my_req = 0, my_batches = 0;                             // service kernel: my request number, batches I am done with
        if (kSvc && wide) {
            my_batches = box->consumed;
            if (lane == 0) {
                box->bs = bs; box->P = ch.P; box->n = remaining;
                box->states = ch.sl | (ch.sm << 10) | (ch.so << 20);
                __threadfence_block();
                st_vol(&box->posted, box->posted + 1);
            }
            __syncwarp();
            my_req = box->posted;
        }

        while (remaining > 0 && !stop) {
            if (wide) {
                int produced = 0;
                bool final_batch = false;
                const uint2 *recs = sm.rec[0];
                if (!kSvc) {
                    if (lane == 0) {
                        if (ch.P >= kFastMinBits) ch.open();
                        while (produced < kSeqBatch && remaining > 0 && ch.P >= kFastMinBits) {
                            sm.rec[0][produced] = ch.step(sm.ll, sm.ml, sm.of);
                            produced++;
                            remaining--;
                            b_stale = true;
                        }
                    }
                    produced = __shfl_sync(kFull, produced, 0);
                    remaining = __shfl_sync(kFull, remaining, 0);
                    final_batch = produced < kSeqBatch;
                }
                else {
                    for (;;) {                                            // the next batch of MY request (older ones are dropped)
                        while ((int32_t) (ld_vol(&box->produced) - my_batches) <= 0) pause_ns(100);
                        __threadfence_block();
                        const uint32_t c = box->count[my_batches & (kSlots - 1)];
                        if (((c >> 8) & 0x7FFFFFu) == (my_req & 0x7FFFFFu)) {
                            produced = (int) (c & 0xFF);
                            final_batch = (c & kChainFinal) != 0;
                            recs = sm.rec[my_batches & (kSlots - 1)];
                            break;
                        }
                        __syncwarp();
                        my_batches++;
                        if (lane == 0) st_vol(&box->consumed, my_batches);
                    }
                    remaining -= produced;
                    if (produced) b_stale = true;
                    if (final_batch && lane == 0) {
                        ch.P = box->end_P;
                        const uint32_t st = box->end_states;
                        ch.sl = st & 1023; ch.sm = (st >> 10) & 1023; ch.so = st >> 20;
                    }
                }
#ifdef LZS_EMU
                if (lane == 0) emu_count_wide(produced);
#endif
                const bool leave = remaining > 0 && final_batch;          // close to the start of the stream: the exact loop takes over
                __syncwarp();
                if (produced) {
                    // ---- every lane: the lengths and the offset code of ITS sequence
                    int32_t my_ll = 0, my_ml = 0;
                    uint32_t my_key = 0;                                  // a new offset, or 0x80000000 | repeated-offset code 0..3
                    if (lane < produced) {
                        const int32_t ps = (int32_t) recs[lane].x;
                        const uint32_t c = recs[lane].y;
                        const uint32_t llc = c & 0xFF, mlc = (c >> 8) & 0xFF, ofc = (c >> 16) & 0xFF;
                        int32_t a = (ps - 57) >> 3;
                        uint64_t v = ld64u(bs + a) << (64 - (ps - a * 8));   // the next unread bit is bit 63
                        const uint32_t ofx = (uint32_t) ((v >> 1) >> (63 - ofc));
                        const int32_t q = ps - (int32_t) ofc;
                        const uint32_t mlb = kMLBits[mlc], llb = kLLBits[llc];
                        a = (q - 57) >> 3;
                        v = ld64u(bs + a) << (64 - (q - a * 8));
                        const uint32_t mlx = (uint32_t) ((v >> 1) >> (63 - mlb));
                        v <<= mlb;
                        const uint32_t llx = (uint32_t) ((v >> 1) >> (63 - llb));
                        my_ml = kMLBase[mlc] + (int32_t) mlx;
                        my_ll = kLLBase[llc] + (int32_t) llx;
                        const uint32_t ofv = (uint32_t) of_base((int) ofc) + ofx;
                        my_key = ofc <= 1 ? (0x80000000u | (ofv + (llc == 0 ? 1u : 0u))) : ofv;
                    }
                    // ---- repeated offsets, in sequence order (:380-408); every lane keeps the history, lane s keeps offset s
                    int32_t my_off = 0;
                    for (int s = 0; s < produced; s++) {
                        const uint32_t k = __shfl_sync(kFull, my_key, s);
                        if (k & 0x80000000u) {
                            const uint32_t o = k & 3;
                            if (o != 0) {
                                int32_t temp = o == 3 ? p0 - 1 : (o == 1 ? p1 : p2);
                                if (temp == 0) temp = 1;
                                if (o != 1) p2 = p1;
                                p1 = p0;
                                p0 = temp;
                            }
                        }
                        else { p2 = p1; p1 = p0; p0 = (int32_t) k; }
                        if (lane == s) my_off = p0;
                    }
                    // ---- positions: prefix sums over the batch
                    const int32_t my_total = my_ll + my_ml;
                    int32_t a = my_ll, t = my_total;
                    for (int o = 1; o < 32; o <<= 1) {
                        const int32_t ua = __shfl_up_sync(kFull, a, o), ut = __shfl_up_sync(kFull, t, o);
                        if (lane >= o) { a += ua; t += ut; }
                    }
                    const int32_t my_opw = opw + t - my_total;            // where my output starts
                    const int32_t my_mop = my_opw + my_ll;                // ... and where my match starts
                    const int32_t my_dl = (a - my_ll) - my_opw;           // literal index in the batch = output position + my_dl
                    // ---- the three checks of the Java loop, first failing sequence first
                    int bad = 0;
                    if (lane < produced) {
                        if (abs0 + (my_opw - oh) + my_total > out_cap) bad = R_OUTPUT_TOO_SMALL;
                        else if (lit_pos + a > lit.size) bad = R_CORRUPTED;
                        else if (abs0 + (my_mop - oh) - my_off < 0) bad = R_CORRUPTED;
                    }
                    const unsigned badm = __ballot_sync(kFull, bad != 0);
                    if (badm) ZFAIL(__shfl_sync(kFull, bad, __ffs((int) badm) - 1), input);
                    const uint8_t *const litb = lit.ptr + lit_pos;
                    if (lane < produced) {                                // ask for the lines now: the steps below then find them in L1
                        if (my_ll) prefetch_l1(litb + (a - my_ll));
                        if (my_mop - my_off < my_mop + my_ml - kRing) prefetch_l1(out_al + (my_mop - my_off));
                    }
                    int f = 0;
                    while (f < produced) {
                        // ---- a step: as many consecutive sequences as end within 64 bytes of the first one's start S and take all
                        // their match bytes from in front of S (then no byte of the step depends on another byte of the step)
                        const int32_t S = __shfl_sync(kFull, my_opw, f);
                        const int32_t endk = my_opw + my_total - S;
                        const bool fits = lane >= f && lane < produced && endk <= 64 && my_off >= endk;
                        const uint32_t run = __ballot_sync(kFull, fits) >> f;
                        const int nfit = __ffs((int) ~run) - 1;           // a sequence has at least 3 bytes: never 32 of them in a step
                        if (nfit == 0) {
                            ring_long_sequence(sm.ring, litb + __shfl_sync(kFull, a - my_ll, f), out_al, flushed, ring_lo, S,
                                               __shfl_sync(kFull, my_ll, f), __shfl_sync(kFull, my_ml, f), __shfl_sync(kFull, my_off, f), lane);
                            f++;
                            continue;
                        }
                        const int32_t E = __shfl_sync(kFull, endk, f + nfit - 1);   // bytes of this step
                        const int32_t rlo = S + E - kRing > ring_lo ? S + E - kRing : ring_lo;   // oldest position the ring still holds
                        // which sequence produces byte j of the step: count the sequence starts at or below j
                        const bool inwin = lane >= f && lane < f + nfit;
                        const int32_t st = my_opw - S;
                        const uint32_t lo = __reduce_or_sync(kFull, (inwin && st < 32) ? 1u << st : 0u);
                        {
                            const int r = f - 1 + __popc(lo & le_mask);
                            const int32_t mop = __shfl_sync(kFull, my_mop, r), dl = __shfl_sync(kFull, my_dl, r), off = __shfl_sync(kFull, my_off, r);
                            const int32_t pos = S + lane;
                            if (lane < E) {
                                uint8_t v;
                                if (pos < mop) v = litb[pos + dl];
                                else {
                                    const int32_t src = pos - off;
                                    v = src >= rlo ? sm.ring[src & kRingMask] : out_al[src];
                                }
                                sm.ring[pos & kRingMask] = v;
                            }
                        }
                        if (E > 32) {
                            const uint32_t hi = __reduce_or_sync(kFull, (inwin && st >= 32) ? 1u << (st - 32) : 0u);
                            const int r = f - 1 + __popc(lo) + __popc(hi & le_mask);
                            const int32_t mop = __shfl_sync(kFull, my_mop, r), dl = __shfl_sync(kFull, my_dl, r), off = __shfl_sync(kFull, my_off, r);
                            const int32_t pos = S + 32 + lane;
                            if (lane + 32 < E) {
                                uint8_t v;
                                if (pos < mop) v = litb[pos + dl];
                                else {
                                    const int32_t src = pos - off;
                                    v = src >= rlo ? sm.ring[src & kRingMask] : out_al[src];
                                }
                                sm.ring[pos & kRingMask] = v;
                            }
                        }
                        __syncwarp();
                        f += nfit;
                        if (S + E - flushed >= kFlush) ring_flush(sm.ring, out_al, flushed, (S + E) & ~15, lane);
                    }
                    lit_pos += __shfl_sync(kFull, a, 31);
                    opw += __shfl_sync(kFull, t, 31);
                    output = abs0 + (opw - oh);
                }
                if (kSvc) {                                               // the slot is free again
                    __syncwarp();
                    my_batches++;
                    if (lane == 0) st_vol(&box->consumed, my_batches);
                }
                if (leave) {
                    if (opw != flushed) ring_flush(sm.ring, out_al, flushed, opw, lane);
                    wide = false;
                    if (lane == 0 && b_stale) {                           // the state BitInputStream.load() is in at this point
                        const int32_t cr = (ch.P - 57) >> 3;
                        b.cur = b.start + cr;
                        b.consumed = cr * 8 + 64 - ch.P;
                        b.bits = ld64u(b.in + b.cur);
                        b.overflow = 0;
                        ll_state = (int) ch.sl; ml_state = (int) ch.sm; of_state = (int) ch.so;
                    }
                }
                continue;
            }
            int produced = 0;
            int fail = 0;
            if (lane == 0) {
                while (produced < kSeqBatch && remaining > 0) {
                    remaining--;
                    br_load(b);
                    if (b.overflow) {
                        if (remaining != 0) fail = R_NOT_ALL_SEQUENCES;
                        stop = true;
                        break;
                    }
                    const uint32_t el = sm.ll[ll_state], em = sm.ml[ml_state], eof = sm.of[of_state];
                    const int ll_code = el >> 24, ml_code = em >> 24, of_code = eof >> 24;
                    const int ll_bits = kLLBits[ll_code], ml_bits = kMLBits[ml_code], of_bits = of_code;
                    int32_t offset = of_base(of_code);
                    if (of_code > 0) { offset += (int32_t) peek_bits(b.consumed, b.bits, of_bits); b.consumed += of_bits; }
                    if (of_code <= 1) {
                        if (ll_code == 0) offset++;
                        if (offset != 0) {
                            int32_t temp = (offset == 3) ? p0 - 1 : (offset == 1 ? p1 : p2);
                            if (temp == 0) temp = 1;
                            if (offset != 1) p2 = p1;
                            p1 = p0;
                            p0 = temp;
                            offset = temp;
                        }
                        else {
                            offset = p0;
                        }
                    }
                    else {
                        p2 = p1; p1 = p0; p0 = offset;
                    }
                    int32_t match_length = kMLBase[ml_code];
                    if (ml_code > 31) { match_length += (int32_t) peek_bits(b.consumed, b.bits, ml_bits); b.consumed += ml_bits; }
                    int32_t lit_length = kLLBase[ll_code];
                    if (ll_code > 15) { lit_length += (int32_t) peek_bits(b.consumed, b.bits, ll_bits); b.consumed += ll_bits; }
                    if (ll_bits + ml_bits + of_bits > 64 - 7 - (9 + 9 + 8)) br_load(b);
                    int nb;
                    nb = (el >> 19) & 0x1F; ll_state = (int) ((el & 0x3FF) + (uint32_t) peek_bits(b.consumed, b.bits, nb)); b.consumed += nb;
                    nb = (em >> 19) & 0x1F; ml_state = (int) ((em & 0x3FF) + (uint32_t) peek_bits(b.consumed, b.bits, nb)); b.consumed += nb;
                    nb = (eof >> 19) & 0x1F; of_state = (int) ((eof & 0x3FF) + (uint32_t) peek_bits(b.consumed, b.bits, nb)); b.consumed += nb;
                    sm.seq_ll[produced] = lit_length;
                    sm.seq_ml[produced] = match_length;
                    sm.seq_of[produced] = offset;
                    produced++;
                }
            }
#ifdef LZS_EMU
            if (lane == 0) emu_count_exact(produced);
#endif
            produced = __shfl_sync(kFull, produced, 0);
            fail = __shfl_sync(kFull, fail, 0);
            remaining = __shfl_sync(kFull, remaining, 0);
            stop = __shfl_sync(kFull, (int) stop, 0) != 0;
            __syncwarp();
            // Every sequence of the batch knows where its literals and its match source lie (prefix sums over the batch), so
            // lane s asks for the lines of sequence s now: the copies below then find them in L1 instead of paying one
            // L2 round trip per sequence (stores do not allocate in L1, and the literal buffer is global scratch).
            {
                int32_t a = lane < produced ? sm.seq_ll[lane] : 0, t = lane < produced ? sm.seq_ll[lane] + sm.seq_ml[lane] : 0;
                const int32_t my_ll = a;
                for (int o = 1; o < 32; o <<= 1) {
                    const int32_t ua = __shfl_up_sync(kFull, a, o), ut = __shfl_up_sync(kFull, t, o);
                    if (lane >= o) { a += ua; t += ut; }
                }
                if (lane < produced) {
                    const int64_t lp = lit_pos + (a - my_ll);                       // literal position of sequence `lane`
                    const int64_t ms = output + (t - sm.seq_ml[lane]) - sm.seq_of[lane];   // its match source
                    if (lit.rle < 0 && my_ll > 0 && lp < lit.size) prefetch_l1(lit.ptr + lp);
                    if (ms >= 0 && ms < output) prefetch_l1(out + ms);
                }
            }
            // execute (all lanes, in order)
            for (int s = 0; s < produced; s++) {
                const int64_t lit_length = sm.seq_ll[s], match_length = sm.seq_ml[s];
                const int64_t offset = sm.seq_of[s];
                const int64_t lit_out_limit = output + lit_length;
                const int64_t match_out_limit = lit_out_limit + match_length;
                ZCHECK(match_out_limit <= out_cap, input, R_OUTPUT_TOO_SMALL);
                const int64_t lit_end = lit_pos + lit_length;
                ZCHECK(lit_end <= lit.size, input, R_CORRUPTED);
                ZCHECK(lit_out_limit - offset >= 0, input, R_CORRUPTED);
                const int64_t total = lit_length + match_length;
                if (total <= 32 && offset >= total && lit.rle < 0) {
                    // the whole sequence in one step: a lane's byte is a literal or a match byte in front of this sequence
                    if (lane < total) {
                        const uint8_t *src = lane < lit_length ? lit.ptr + (lit_pos + lane) : out + (output + lane - offset);
                        out[output + lane] = *src;
                    }
                }
                else {
                    copy_literals(out + output, lit, lit_pos, lit_length, lane);
                    __syncwarp();
                    warp_match_copy(out + lit_out_limit, offset, match_length, lane);
                }
                __syncwarp();
                output = match_out_limit;
                lit_pos = lit_end;
            }
            if (fail) ZFAIL(fail, input);
            __syncwarp();
        }
        if (wide && opw != flushed) ring_flush(sm.ring, out_al, flushed, opw, lane);   // the block ended on the wide path
        fs.prev[0] = __shfl_sync(kFull, p0, 0);
        fs.prev[1] = __shfl_sync(kFull, p1, 0);
        fs.prev[2] = __shfl_sync(kFull, p2, 0);
    }
    // copyLastLiteral :518-525
    const int64_t last = lit.size - lit_pos;
    ZCHECK(output + last <= out_cap, input, R_OUTPUT_TOO_SMALL);
    copy_literals(out + output, lit, lit_pos, last, lane);
    output += last;
    __syncwarp();
    return output - out_pos;
}

// ZstdFrameDecompressor.decompress :135-210 for one input.  Returns output size or -1.
template <bool kSvc>
__device__ int64_t decode_input(WarpSmem &sm, ChainBox *box, const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, uint8_t *lit_scratch, Ctl &ctl, int lane)
{
    if (out_cap == 0) return 0;
    int64_t input = 0, output = 0;
    FrameState fs;
    fs.huf_log = -1;   // a fresh decompressor per call (the Java object would keep its table across calls)
    while (input < in_len) {
        fs.prev[0] = 1; fs.prev[1] = 4; fs.prev[2] = 8;
        fs.ll_log = fs.of_log = fs.ml_log = -1;
        const int64_t output_start = output;
        // verifyMagic :949-962
        ZCHECK(in_len - input >= 4, input, R_NOT_ENOUGH_INPUT);
        const uint32_t magic = ld32u(in + input);
        if (magic != 0xFD2FB528u) ZFAIL(magic == 0xFD2FB527u ? R_V07_FORMAT : R_BAD_MAGIC, input);
        input += 4;
        // readFrameHeader :860-940
        const int64_t fh_start = input;
        ZCHECK(input < in_len, input, R_NOT_ENOUGH_INPUT);
        const int fhd = in[input++];
        const bool single_segment = (fhd & 0x20) != 0;
        const int dict_desc = fhd & 3, cs_desc = fhd >> 6;
        const int header_size = 1 + (single_segment ? 0 : 1) + (dict_desc == 0 ? 0 : (1 << (dict_desc - 1))) +
                                (cs_desc == 0 ? (single_segment ? 1 : 0) : (1 << cs_desc));
        ZCHECK(header_size <= in_len - fh_start, input, R_NOT_ENOUGH_INPUT);
        int32_t window_size = -1;
        if (!single_segment) {
            const int wd = in[input++];
            const int exponent = wd >> 3, mantissa = wd & 7;
            const int32_t base = (int32_t) (1u << ((10 + exponent) & 31));
            window_size = (int32_t) ((uint32_t) base + (uint32_t) (base / 8) * (uint32_t) mantissa);
        }
        if (dict_desc != 0) { input += (1 << (dict_desc - 1)); ZFAIL(R_DICTIONARY, input); }
        input = fh_start + header_size;   // content size field is not needed for decoding
        const bool has_checksum = (fhd & 4) != 0;

        bool last_block;
        do {
            ZCHECK(input + 3 <= in_len, input, R_NOT_ENOUGH_INPUT);
            const uint32_t header = ld16u(in + input) | ((uint32_t) in[input + 2] << 16);
            input += 3;
            last_block = (header & 1) != 0;
            const int block_type = (header >> 1) & 3;
            const int32_t block_size = (int32_t) ((header >> 3) & 0x1FFFFF);
            int64_t decoded;
            if (block_type == 0) {
                ZCHECK(input + block_size <= in_len, input, R_NOT_ENOUGH_INPUT);
                ZCHECK(output + block_size <= out_cap, input, R_OUTPUT_TOO_SMALL);
                warp_copy(out + output, in + input, block_size, lane);
                decoded = block_size;
                input += block_size;
            }
            else if (block_type == 1) {
                ZCHECK(input + 1 <= in_len, input, R_NOT_ENOUGH_INPUT);
                ZCHECK(output + block_size <= out_cap, input, R_OUTPUT_TOO_SMALL);
                const uint8_t v = in[input];
                for (int64_t i = lane; i < block_size; i += 32) out[output + i] = v;
                decoded = block_size;
                input += 1;
            }
            else if (block_type == 2) {
                ZCHECK(input + block_size <= in_len, input, R_NOT_ENOUGH_INPUT);
                // does anything follow this block in the input (another block, or another frame after the checksum)?
                const bool more_follows = !last_block || input + block_size + (has_checksum ? 4 : 0) < in_len;
                decoded = decode_compressed_block<kSvc>(sm, box, fs, in, input, block_size, out, output, out_cap, window_size, lit_scratch, more_follows,
                                                  ctl, lane);
                if (decoded < 0) return -1;
                input += block_size;
            }
            else {
                ZFAIL(R_INVALID_BLOCK_TYPE, input);
            }
            output += decoded;
            __syncwarp();
        }
        while (!last_block);

        if (has_checksum) {
            __syncwarp();
#ifndef LZS_EMU
            const uint64_t hash = xxh64_group4(out + output_start, lane < 4 ? output - output_start : 0, 0, lane & 3, 0xFu << (lane & ~3));
#else
            const uint64_t hash = emu_xxh64(out + output_start, output - output_start);   // host emulation: scalar, no group shuffles
#endif
            const uint32_t h32 = (uint32_t) __shfl_sync(kFull, hash, 0);
            ZCHECK(input + 4 <= in_len, input, R_NOT_ENOUGH_INPUT);
            if (ld32u(in + input) != h32) ZFAIL(R_BAD_CHECKSUM, input);
            input += 4;
        }
    }
    return output;
}

#ifndef LZS_EMU
template <int kCtasPerSm>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) zstd_decompress_kernel(AccBatch b, uint8_t *scratch, int64_t scratch_per_warp)
{
    extern __shared__ __align__(16) uint8_t zsmem[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    WarpSmem &sm = *reinterpret_cast<WarpSmem *>(zsmem + (size_t) warp * sizeof(WarpSmem));
    uint8_t *lit_scratch = scratch + ((int64_t) blockIdx.x * kWarpsPerCta + warp) * scratch_per_warp;
    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;
        Ctl ctl;
        ctl.reason = 0; ctl.err_off = 0;
        int64_t r = decode_input<false>(sm, nullptr, b.src + b.src_off[idx], b.src_len[idx], b.dst + b.dst_off[idx], b.dst_cap[idx], lit_scratch, ctl, lane);
        if (lane == 0) {
            if (r >= 0) { b.out_len[idx] = r; b.status[idx] = 0; }
            else { b.out_len[idx] = ctl.err_off; b.status[idx] = ACC_STATUS(ACC_E_MALFORMED, ctl.reason); }
        }
        __syncwarp();
    }
}

// the service kernel: kWorkers worker warps (one input each at a time, claimed from the work counter) + the chain warp
template <int kWorkers, int kCtasPerSm>
__global__ void __launch_bounds__((kWorkers + 1) * 32, kCtasPerSm) zstd_decompress_svc_kernel(AccBatch b, uint8_t *scratch, int64_t scratch_per_warp)
{
    extern __shared__ __align__(16) uint8_t zsmem[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    WarpSmem *const sms = reinterpret_cast<WarpSmem *>(zsmem);
    ChainBox *const boxes = reinterpret_cast<ChainBox *>(zsmem + (size_t) kWorkers * sizeof(WarpSmem));
    uint32_t *const workers_done = reinterpret_cast<uint32_t *>(boxes + kWorkers);
    if (threadIdx.x < kWorkers) {
        ChainBox &x = boxes[threadIdx.x];
        x.posted = 0; x.consumed = 0; x.produced = 0; x.count[0] = 0; x.count[1] = 0;
    }
    if (threadIdx.x == 0) *workers_done = 0;
    __syncthreads();
    if (warp == kWorkers) { chain_warp<kWorkers>(boxes, sms, workers_done, lane); return; }
    WarpSmem &sm = sms[warp];
    uint8_t *lit_scratch = scratch + ((int64_t) blockIdx.x * kWorkers + warp) * scratch_per_warp;
    for (;;) {
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(b.work_counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if ((int64_t) idx >= b.n) break;
        Ctl ctl;
        ctl.reason = 0; ctl.err_off = 0;
        int64_t r = decode_input<true>(sm, boxes + warp, b.src + b.src_off[idx], b.src_len[idx], b.dst + b.dst_off[idx], b.dst_cap[idx], lit_scratch, ctl, lane);
        if (r < 0) chain_cancel(boxes + warp, lane);   // the chain lane may still be walking this input's tables: make it stop before they are reused
        if (lane == 0) {
            if (r >= 0) { b.out_len[idx] = r; b.status[idx] = 0; }
            else { b.out_len[idx] = ctl.err_off; b.status[idx] = ACC_STATUS(ACC_E_MALFORMED, ctl.reason); }
        }
        __syncwarp();
    }
    if (lane == 0) atomicAdd(workers_done, 1u);
}
#endif  // LZS_EMU
}  // namespace

static constexpr int64_t kZstdDecScratchPerWarp = zs::kMaxBlock + 256 + kHufBytes + kFseBytes;   // literals | parked Huffman table | parked FSE tables

constexpr int kZstdDecMaxCtasPerSm = 7;   // the scratch is sized for 7 x 4 = 28 decoding warps per SM (the most any kernel shape keeps resident)

int64_t acc_zstd_dec_grid(int sm_count) { return (int64_t) sm_count * kZstdDecMaxCtasPerSm; }

int64_t acc_zstd_dec_scratch_bytes(int sm_count) { return acc_zstd_dec_grid(sm_count) * kWarpsPerCta * kZstdDecScratchPerWarp; }

#ifndef LZS_EMU
template <int kCtasPerSm>
static void launch_zstd_decompress(const AccBatch &b, int sm_count, cudaStream_t st, void *scratch)
{
    const int smem = kWarpsPerCta * (int) sizeof(WarpSmem);
    cudaFuncSetAttribute(zstd_decompress_kernel<kCtasPerSm>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int64_t ctas = (b.n + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t max_ctas = (int64_t) sm_count * kCtasPerSm;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    zstd_decompress_kernel<kCtasPerSm><<<(unsigned) ctas, kWarpsPerCta * 32, smem, st>>>(b, (uint8_t *) scratch, kZstdDecScratchPerWarp);
}

template <int kWorkers, int kCtasPerSm>
static void launch_zstd_decompress_svc(const AccBatch &b, int sm_count, cudaStream_t st, void *scratch)
{
    static_assert(kWorkers * kCtasPerSm <= kZstdDecMaxCtasPerSm * kWarpsPerCta, "scratch is sized for 28 decoding warps per SM");
    const int smem = kWorkers * (int) sizeof(WarpSmem) + kWorkers * (int) sizeof(ChainBox) + 16;
    cudaFuncSetAttribute(zstd_decompress_svc_kernel<kWorkers, kCtasPerSm>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int64_t ctas = (b.n + kWorkers - 1) / kWorkers;
    const int64_t max_ctas = (int64_t) sm_count * kCtasPerSm;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    zstd_decompress_svc_kernel<kWorkers, kCtasPerSm><<<(unsigned) ctas, (kWorkers + 1) * 32, smem, st>>>(b, (uint8_t *) scratch, kZstdDecScratchPerWarp);
}

// ctas_per_sm (acc_set_tuning key 0): 0 = automatic -- the service kernel, 13 workers + chain warp x 2 CTAs per SM when the batch
// fills that shape, 7 + 1 x 3 for smaller batches (more CTAs to spread over the SMs); 173 / 232 / 371 = the service kernel as
// 7+1 x 3, 13+1 x 2, 27+1 x 1; 5 = the warp-per-input kernel (the state walk inline on lane 0), 5 CTAs of 4 warps per SM.
// Measured on 32,768 128 KiB Silesia blocks: 52.5 / 55.7 / 56.2 GiB/s for the three service shapes, 44-47 for the warp-per-input kernel.
void acc_launch_zstd_decompress(const AccBatch &b, int sm_count, int ctas_per_sm, cudaStream_t st, void *scratch, int64_t scratch_bytes)
{
    (void) scratch_bytes;
    if (ctas_per_sm == 0) ctas_per_sm = b.n >= (int64_t) sm_count * 26 ? 232 : 173;
    if (ctas_per_sm == 232) launch_zstd_decompress_svc<13, 2>(b, sm_count, st, scratch);
    else if (ctas_per_sm == 371) launch_zstd_decompress_svc<27, 1>(b, sm_count, st, scratch);
    else if (ctas_per_sm == 5) launch_zstd_decompress<5>(b, sm_count, st, scratch);
    else launch_zstd_decompress_svc<7, 3>(b, sm_count, st, scratch);
}
#endif  // LZS_EMU
