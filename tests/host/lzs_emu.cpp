// lzs_emu.cpp -- runs the two-phase LZ77 decode engine (aircompressor_b200/csrc/lz_stream.cuh with the LZ4 and Snappy parse
// sides) on the CPU: OS threads play the lanes.  The decoded blocks, lengths and status words are written to a file that
// tests/test_stream_engine_emu.py compares with the oracle.  TEST INFRASTRUCTURE: nothing here ships.
//
//   lzs_emu <in-file> <out-file>
//   in-file : int32 codec (0 lz4, 1 snappy), int32 n, then per block { int64 in_len, int64 out_cap, int32 in_misalign,
//             int32 out_misalign, in_len bytes }
//   out-file: per block { int64 out_len, int32 status, out_cap + 64 bytes (the 64 guard bytes must stay 0xA5) }
#define LZS_EMU 1
#include "cuda_emu.h"

#include <thread>
#include <vector>

thread_local EmuWarp *t_warp = nullptr;
thread_local int t_lane = 0;

#include "../../aircompressor_b200/csrc/lz4_stream.cuh"
#include "../../aircompressor_b200/csrc/snappy_decode.cuh"

constexpr int kWarps = 3;

template <class Codec>
static void run(const AccBatch &b, int lanes_in_use)
{
    static lzs::WarpSmem sm[kWarps];
    EmuWarp warps[kWarps];
    for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int w = 0; w < kWarps; w++)
        for (int l = 0; l < 32; l++)
            th.emplace_back([&, w, l] {
                t_warp = &warps[w];
                t_lane = l;
                lzs::run_warp<Codec>(b, sm[w], l, lanes_in_use);
            });
    for (auto &t : th) t.join();
}

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t codec = 0, n = 0;
    if (fread(&codec, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1) return 2;
    std::vector<int64_t> src_off(n), src_len(n), dst_off(n), dst_cap(n), out_len(n, -12345);
    std::vector<int32_t> status(n, -777);
    std::vector<std::vector<uint8_t>> ins(n);
    std::vector<int32_t> imis(n), omis(n);
    int64_t sp = 64, dp = 64;
    for (int i = 0; i < n; i++) {
        int64_t hdr[2];
        int32_t mis[2];
        if (fread(hdr, 8, 2, f) != 2 || fread(mis, 4, 2, f) != 2) return 2;
        src_len[i] = hdr[0]; dst_cap[i] = hdr[1]; imis[i] = mis[0]; omis[i] = mis[1];
        ins[i].resize(hdr[0]);
        if (hdr[0] && fread(ins[i].data(), 1, hdr[0], f) != (size_t) hdr[0]) return 2;
        sp = ((sp + 15) & ~15LL) + imis[i];
        src_off[i] = sp; sp += src_len[i];
        dp = ((dp + 15) & ~15LL) + omis[i];
        dst_off[i] = dp; dp += dst_cap[i] + 64;
    }
    fclose(f);
    uint8_t *src = (uint8_t *) aligned_alloc(4096, (size_t) ((sp + 64 + 4095) & ~4095LL));
    uint8_t *dst = (uint8_t *) aligned_alloc(4096, (size_t) ((dp + 64 + 4095) & ~4095LL));
    memset(src, 0x5A, sp + 64);
    memset(dst, 0xA5, dp + 64);
    for (int i = 0; i < n; i++) if (src_len[i]) memcpy(src + src_off[i], ins[i].data(), src_len[i]);
    unsigned int counter = 0;
    AccBatch b;
    b.src = src; b.src_off = src_off.data(); b.src_len = src_len.data();
    b.dst = dst; b.dst_off = dst_off.data(); b.dst_cap = dst_cap.data();
    b.out_len = out_len.data(); b.status = status.data(); b.n = n; b.work_counter = &counter;

    // lanes in use per warp as the launcher computes them (every lane claims blocks until the batch is exhausted)
    int lanes = (n + kWarps - 1) / kWarps;
    if (lanes > 32) lanes = 32;
    if (lanes < 1) lanes = 1;
    if (getenv("LZS_EMU_LANES")) lanes = atoi(getenv("LZS_EMU_LANES"));
    if (codec == 0) run<lz4v1::Lz4Stream>(b, lanes); else run<snappydec::SnappyStream>(b, lanes);

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 2;
    for (int i = 0; i < n; i++) {
        fwrite(&out_len[i], 8, 1, o);
        fwrite(&status[i], 4, 1, o);
        fwrite(dst + dst_off[i], 1, dst_cap[i] + 64, o);
    }
    fclose(o);
    fprintf(stderr, "lzs_emu: %d blocks, %d lanes per warp in use\n", n, lanes);
    return 0;
}
