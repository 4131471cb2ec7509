// lzs_emu.cpp -- runs the record path of the LZ4 / Snappy decoders (aircompressor_b200/csrc/lz_records.cuh with both parse
// sides, and the step decoders it resumes) on the CPU: OS threads play the lanes.  Stage 1 = the parse kernel (independent
// lanes), stage 2 = the execute kernel (warps of 32 threads, barriers as __syncwarp, an exchange array as shuffles).  The
// decoded blocks, lengths and status words are written to a file that tests/test_record_engine_emu.py compares with the
// oracle.  TEST INFRASTRUCTURE: nothing here ships.
//
//   lzs_emu <in-file> <out-file> [row]
//   in-file : int32 codec (0 lz4, 1 snappy), int32 n, then per block { int64 in_len, int64 out_cap, int32 in_misalign,
//             int32 out_misalign, in_len bytes }
//   out-file: per block { int64 out_len, int32 status, out_cap + 64 bytes (the 64 guard bytes must stay 0xA5) }
#define LZS_EMU 1
#include "cuda_emu.h"

#include <thread>
#include <vector>

thread_local EmuWarp *t_warp = nullptr;
thread_local int t_lane = 0;

#include "../../aircompressor_b200/csrc/lz4_records.cuh"
#include "../../aircompressor_b200/csrc/snappy_decode.cuh"

constexpr int kParseLanes = 8, kWarps = 3;

template <class Codec>
static void run(AccBatch b, int row, long *n_records)
{
    std::vector<uint2> recs((size_t) b.n * row);
    std::vector<lzs::RecHeader> hdrs(b.n);
    unsigned int c1 = 0, c2 = 0;
    {   // stage 1: parse
        b.work_counter = &c1;
        static uint8_t win[kParseLanes][lzs::kWinStride] __attribute__((aligned(16)));
        EmuWarp pw;                                       // the parse lanes run their rounds in lockstep like the lanes of a warp
        pw.nlanes = kParseLanes;
        pthread_barrier_init(&pw.bar, nullptr, kParseLanes);
        std::vector<std::thread> th;
        for (int l = 0; l < kParseLanes; l++) th.emplace_back([&, l] { t_warp = &pw; t_lane = l; lzs::parse_lane<Codec>(b, win[l], recs.data(), hdrs.data(), row); });
        for (auto &t : th) t.join();
    }
    for (auto &h : hdrs) *n_records += h.n_rec;
    {   // stage 2: execute
        b.work_counter = &c2;
        EmuWarp warps[kWarps];
        static uint8_t rings[kWarps][lzs::kOutRing] __attribute__((aligned(16)));
        for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
        std::vector<std::thread> th;
        for (int w = 0; w < kWarps; w++)
            for (int l = 0; l < 32; l++)
                th.emplace_back([&, w, l] { t_warp = &warps[w]; t_lane = l; lzs::execute_warp<Codec>(b, recs.data(), hdrs.data(), row, rings[w], l); });
        for (auto &t : th) t.join();
    }
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const int row = argc > 3 ? atoi(argv[3]) : 8192;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t codec = 0, n = 0;
    if (fread(&codec, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1) return 2;
    std::vector<int64_t> src_off(n), src_len(n), dst_off(n), dst_cap(n), out_len(n, -12345);
    std::vector<int32_t> status(n, -777);
    std::vector<std::vector<uint8_t>> ins(n);
    int64_t sp = 64, dp = 64;
    for (int i = 0; i < n; i++) {
        int64_t hdr[2];
        int32_t mis[2];
        if (fread(hdr, 8, 2, f) != 2 || fread(mis, 4, 2, f) != 2) return 2;
        src_len[i] = hdr[0]; dst_cap[i] = hdr[1];
        ins[i].resize(hdr[0]);
        if (hdr[0] && fread(ins[i].data(), 1, hdr[0], f) != (size_t) hdr[0]) return 2;
        sp = ((sp + 31) & ~31LL) + mis[0];
        src_off[i] = sp; sp += src_len[i];
        dp = ((dp + 15) & ~15LL) + mis[1];
        dst_off[i] = dp; dp += dst_cap[i] + 64;
    }
    fclose(f);
    uint8_t *src = (uint8_t *) aligned_alloc(4096, (size_t) ((sp + 64 + 4095) & ~4095LL));
    uint8_t *dst = (uint8_t *) aligned_alloc(4096, (size_t) ((dp + 64 + 4095) & ~4095LL));
    memset(src, 0x5A, sp + 64);
    memset(dst, 0xA5, dp + 64);
    for (int i = 0; i < n; i++) if (src_len[i]) memcpy(src + src_off[i], ins[i].data(), src_len[i]);
    AccBatch b;
    b.src = src; b.src_off = src_off.data(); b.src_len = src_len.data();
    b.dst = dst; b.dst_off = dst_off.data(); b.dst_cap = dst_cap.data();
    b.out_len = out_len.data(); b.status = status.data(); b.n = n; b.work_counter = nullptr;
    long n_records = 0;
    if (codec == 0) run<lz4v1::Lz4Records>(b, row, &n_records); else run<snappydec::SnappyRecords>(b, row, &n_records);
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 2;
    for (int i = 0; i < n; i++) {
        fwrite(&out_len[i], 8, 1, o);
        fwrite(&status[i], 4, 1, o);
        fwrite(dst + dst_off[i], 1, dst_cap[i] + 64, o);
    }
    fclose(o);
    fprintf(stderr, "lzs_emu: %d blocks, rows of %d, %ld records\n", n, row, n_records);
    return 0;
}
