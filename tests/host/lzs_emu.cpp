// lzs_emu.cpp -- runs the streaming LZ77 decode engine (aircompressor_b200/csrc/lz_stream.cuh with the LZ4 and Snappy parse
// sides) on the CPU: OS threads play the lanes, a DMA thread lands the "bulk copies" late and out of order after
// poisoning their destination.  The decoded blocks, lengths and status words are written to a file that
// tests/test_stream_engine_emu.py compares with the oracle.  TEST INFRASTRUCTURE: nothing here ships.
//
//   lzs_emu <in-file> <out-file>
//   in-file : int32 codec (0 lz4, 1 snappy), int32 n, then per block { int64 in_len, int64 out_cap, int32 in_misalign,
//             int32 out_misalign, in_len bytes }
//   out-file: per block { int64 out_len, int32 status, out_cap + 64 bytes (the 64 guard bytes must stay 0xA5) }
#define LZS_EMU 1
#include "cuda_emu.h"

#include <mutex>
#include <random>
#include <thread>
#include <vector>

thread_local EmuWarp *t_warp = nullptr;
thread_local int t_lane = 0;

#include "../../aircompressor_b200/csrc/lz4_stream.cuh"
#include "../../aircompressor_b200/csrc/snappy_decode.cuh"

// ---- the DMA thread ------------------------------------------------------------------------------------------------
namespace {
struct Copy { void *dst; const void *src; uint32_t bytes; unsigned long long *bar; };
std::mutex g_mu;
std::vector<Copy> g_pending;
std::atomic<bool> g_stop{false};
std::atomic<long> g_copies{0};

void dma_main()
{
    std::mt19937 rng(12345);
    for (;;) {
        Copy c;
        bool have = false;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            if (!g_pending.empty()) {
                const size_t i = rng() % g_pending.size();     // any order
                c = g_pending[i];
                g_pending[i] = g_pending.back();
                g_pending.pop_back();
                have = true;
            }
        }
        if (!have) {
            if (g_stop.load()) return;
            sched_yield();
            continue;
        }
        for (unsigned k = rng() % 200; k; k--) sched_yield();  // late
        memcpy(c.dst, c.src, c.bytes);
        __atomic_fetch_add(c.bar, 1ull, __ATOMIC_RELEASE);
        g_copies++;
    }
}
}  // namespace

namespace lzs {
void emu_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    if (((uintptr_t) dst & 15) || ((uintptr_t) src & 15) || (bytes & 15) || bytes == 0 || bytes > (uint32_t) kChunk) {
        fprintf(stderr, "emu_bulk_load: bad alignment / size (%p %p %u)\n", dst, src, bytes);
        abort();
    }
    memset(dst, 0xEE, bytes);     // the destination is undefined until the copy has landed
    std::lock_guard<std::mutex> lk(g_mu);
    g_pending.push_back({dst, src, bytes, bar});
}
}  // namespace lzs

std::atomic<long> g_seq_records{0}, g_seq_bytes{0}, g_fallbacks{0}, g_fallback_out{0};
namespace lzs {
void emu_count_record(uint32_t z, uint32_t w, uint32_t x, uint32_t y)
{
    if (z) { g_seq_records++; g_seq_bytes += (z & 0xfff) + (z >> 12); }
    else if (w == kRecFallback) { g_fallbacks++; if (x != kFallbackWhole) g_fallback_out += y; }
}
}  // namespace lzs

constexpr int kSlots = 3;

template <class Codec>
static void run(const AccBatch &b)
{
    static lzs::Slot slots[kSlots];
    for (int i = 0; i < kSlots; i++) lzs::init_slot(slots[i]);
    EmuWarp warps[kSlots + 1];
    for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int w = 0; w <= kSlots; w++)
        for (int l = 0; l < 32; l++)
            th.emplace_back([&, w, l] {
                t_warp = &warps[w];
                t_lane = l;
                lzs::run_warp<Codec, kSlots>(b, slots, w, l);
            });
    for (auto &t : th) t.join();
    for (int i = 0; i < kSlots; i++)
        if (slots[i].abort) { fprintf(stderr, "slot %d aborted (watchdog)\n", i); exit(3); }
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

    std::thread dma(dma_main);
    if (codec == 0) run<lz4v1::Lz4Stream>(b); else run<snappydec::SnappyStream>(b);
    g_stop = true;
    dma.join();

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 2;
    for (int i = 0; i < n; i++) {
        fwrite(&out_len[i], 8, 1, o);
        fwrite(&status[i], 4, 1, o);
        fwrite(dst + dst_off[i], 1, dst_cap[i] + 64, o);
    }
    fclose(o);
    fprintf(stderr, "lzs_emu: %d blocks, %ld bulk copies, %ld sequence records (%ld bytes), %ld hand-overs to the general path (after %ld bytes)\n", n, g_copies.load(),
            g_seq_records.load(), g_seq_bytes.load(), g_fallbacks.load(), g_fallback_out.load());
    return 0;
}
