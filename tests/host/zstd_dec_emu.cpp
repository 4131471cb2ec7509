// zstd_dec_emu.cpp -- runs the Zstandard decode kernel's device code (aircompressor_b200/csrc/zstd_dec.cu: frame and block
// headers, Huffman literals, FSE tables, the wide sequence path with its shared-memory output ring, the exact loop it hands
// over to near the start of a bit stream) on the CPU: 32 OS threads play the lanes of a warp, barriers are __syncwarp, an
// exchange array carries shuffles and ballots.  tests/test_zstd_dec_emu.py compares bytes, lengths, status words and error
// offsets with the oracle.  TEST INFRASTRUCTURE: nothing here ships; the GPU parity tests remain the gate for the kernel.
//
//   zstd_dec_emu <in-file> <out-file> [svc]      svc: the service kernel's roles (kWorkers worker warps + the chain warp of a CTA)
//   in-file : int32 0, int32 n, then per input { int64 in_len, int64 out_cap, int32 in_misalign, int32 out_misalign, in_len bytes }
//   out-file: per input { int64 out_len, int32 status, out_cap + 64 bytes (the 64 guard bytes must stay 0xA5) }
#define LZS_EMU 1
#include "cuda_emu.h"

#include <thread>
#include <vector>

thread_local EmuWarp *t_warp = nullptr;
thread_local int t_lane = 0;

// XXH64 of a buffer, scalar (the frame checksum; the kernel's 4-lane version needs group shuffles the emulation does not have)
static uint64_t emu_xxh64(const uint8_t *p, int64_t len)
{
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    auto rd64 = [](const uint8_t *q) { uint64_t v; memcpy(&v, q, 8); return v; };
    auto rd32 = [](const uint8_t *q) { uint32_t v; memcpy(&v, q, 4); return v; };
    auto rotl = [](uint64_t v, int r) { return (v << r) | (v >> (64 - r)); };
    auto round = [&](uint64_t acc, uint64_t v) { return rotl(acc + v * P2, 31) * P1; };
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        for (; p + 32 <= end; p += 32) { v1 = round(v1, rd64(p)); v2 = round(v2, rd64(p + 8)); v3 = round(v3, rd64(p + 16)); v4 = round(v4, rd64(p + 24)); }
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        for (uint64_t v : {v1, v2, v3, v4}) h = (h ^ round(0, v)) * P1 + P4;
    }
    else h = P5;
    h += (uint64_t) len;
    for (; p + 8 <= end; p += 8) h = rotl(h ^ round(0, rd64(p)), 27) * P1 + P4;
    if (p + 4 <= end) { h = rotl(h ^ (rd32(p) * P1), 23) * P2 + P3; p += 4; }
    for (; p < end; p++) h = rotl(h ^ (*p * P5), 11) * P1;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

static std::atomic<long> g_wide{0}, g_exact{0};   // sequences decoded by the wide path / by the exact loop
static void emu_count_wide(int n) { g_wide += n; }
static void emu_count_exact(int n) { g_exact += n; }

#include "../../aircompressor_b200/csrc/zstd_dec.cu"

constexpr int kWorkers = 2;                         // worker warps of the emulated CTA


int main(int argc, char **argv)
{
    if (argc < 3) return 2;
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

    const bool svc = argc > 3 && !strcmp(argv[3], "svc");
    std::atomic<int> next{0};
    constexpr int kEmuWarps = kWorkers + 1;             // worker warps (+ the chain warp in svc mode)
    EmuWarp warps[kEmuWarps];
    static WarpSmem smem[kWorkers] __attribute__((aligned(16)));
    static ChainBox boxes[kWorkers];
    static uint32_t workers_done = 0;
    memset(boxes, 0, sizeof(boxes));
    std::vector<std::vector<uint8_t>> scratch(kWorkers, std::vector<uint8_t>((size_t) kZstdDecScratchPerWarp + 64));
    for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int w = 0; w < kWorkers; w++)
        for (int l = 0; l < 32; l++)
            th.emplace_back([&, w, l] {
                t_warp = &warps[w]; t_lane = l;
                uint8_t *lit_scratch = (uint8_t *) (((uintptr_t) scratch[w].data() + 15) & ~(uintptr_t) 15);
                for (;;) {
                    int idx = 0;
                    if (l == 0) idx = next.fetch_add(1);
                    idx = __shfl_sync(kFull, idx, 0);
                    if (idx >= n) break;
                    Ctl ctl;
                    ctl.reason = 0; ctl.err_off = 0;
                    int64_t r = svc ? decode_input<true>(smem[w], boxes + w, src + src_off[idx], src_len[idx], dst + dst_off[idx], dst_cap[idx], lit_scratch, ctl, l)
                                    : decode_input<false>(smem[w], nullptr, src + src_off[idx], src_len[idx], dst + dst_off[idx], dst_cap[idx], lit_scratch, ctl, l);
                    if (l == 0) {
                        if (r >= 0) { out_len[idx] = r; status[idx] = 0; }
                        else { out_len[idx] = ctl.err_off; status[idx] = ACC_STATUS(ACC_E_MALFORMED, ctl.reason); }
                    }
                    __syncwarp();
                }
                if (l == 0) __atomic_fetch_add(&workers_done, 1u, __ATOMIC_SEQ_CST);
            });
    if (svc)
        for (int l = 0; l < 32; l++)
            th.emplace_back([&, l] { t_warp = &warps[kWorkers]; t_lane = l; chain_warp<kWorkers>(boxes, smem, &workers_done, l); });
    for (auto &t : th) t.join();

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 2;
    for (int i = 0; i < n; i++) {
        fwrite(&out_len[i], 8, 1, o);
        fwrite(&status[i], 4, 1, o);
        fwrite(dst + dst_off[i], 1, dst_cap[i] + 64, o);
    }
    fclose(o);
    fprintf(stderr, "zstd_dec_emu: %d inputs, %ld sequences on the wide path, %ld in the exact loop\n", n, g_wide.load(), g_exact.load());
    return 0;
}
