// xxh64.cu -- batched one-shot XXH64 for sm_100a.
//
// Replaces the reference's XxHash64.hash(seed, base, address, length) (zstd/XxHash64.java:182-249) and
// the public one-shot XxHash64JavaHasher.hash (xxhash/XxHash64JavaHasher.java:73-124); the native
// binding it stands in for is XXH64(input, length, seed) (xxhash/XxHash64Bindings.java:32-34).
// XXH64 has exactly four independent accumulator chains, so one buffer is served by four lanes (one
// lane per accumulator, 8 bytes of every 32-byte stripe each) and a warp hashes eight buffers at once.
#include "acc_device.cuh"
#include "xxh64_device.cuh"

namespace {

__global__ void __launch_bounds__(256) xxh64_kernel(AccBatch b, uint64_t seed)
{
    const int lane = lane_id();
    const int sub = lane & 3;           // accumulator index
    const int grp = lane >> 2;          // buffer slot inside the warp
    const int64_t warp_global = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps_total = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const unsigned gmask = 0xfu << (grp * 4);

    for (int64_t base = warp_global * 8; base < b.n; base += warps_total * 8) {
        const int64_t idx = base + grp;
        const bool active = idx < b.n;
        const uint8_t *in = active ? b.src + b.src_off[idx] : nullptr;
        const int64_t len = active ? b.src_len[idx] : 0;
        uint64_t h = xxh64_group4(in, len, seed, sub, gmask);
        if (active && sub == 0) {
            b.out_len[idx] = (int64_t) h;
            if (b.status) b.status[idx] = 0;
        }
    }
}

}  // namespace

void acc_launch_xxh64(const AccBatch &b, uint64_t seed, int sm_count, cudaStream_t st)
{
    int64_t warps = (b.n + 7) / 8;
    int64_t ctas = (warps + 7) / 8;
    int64_t max_ctas = (int64_t) sm_count * 8;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    xxh64_kernel<<<(unsigned) ctas, 256, 0, st>>>(b, seed);
}
