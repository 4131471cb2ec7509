// lz4_tpb.cu -- LZ4 block decode kernel, one thread per block (see lz4_tpb.cuh for the decoder itself).
// Blocks the optimistic per-thread decoder declines (malformed streams, capacity corner cases, very large blocks) are
// re-decoded by the whole warp with the exact decoder of lz4_decode_v1.cuh, so the kernel's results are bit-exact with
// Lz4RawDecompressor.java:35-198 in every case.
#include "acc_device.cuh"
#include "lz4_decode_v1.cuh"
#include "lz4_tpb.cuh"

__device__ unsigned long long g_tpb_stats[4];   // blocks, fallbacks (debug counters)

namespace {

constexpr int kTpbThreads = 128;

__global__ void __launch_bounds__(kTpbThreads) lz4_decompress_tpb_kernel(AccBatch b)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t) gridDim.x * blockDim.x;
    for (int64_t base = t - lane; base < b.n; base += nthreads) {   // warp-uniform: 32 consecutive blocks per warp step
        const int64_t idx = base + lane;
        bool fb = false;
        if (idx < b.n) {
            const int64_t in_len = b.src_len[idx], cap = b.dst_cap[idx];
            if (in_len < 1 || in_len > (1 << 24) || cap < 1 || cap > (1 << 30)) fb = true;
            else {
                uint32_t olen = 0;
                const int r = lz4tpb::decode_block(b.src + b.src_off[idx], (uint32_t) in_len, b.dst + b.dst_off[idx], (uint32_t) cap, &olen);
                if (r == lz4tpb::kOk) { b.out_len[idx] = olen; b.status[idx] = 0; }
                else fb = true;
            }
        }
        unsigned m = __ballot_sync(kFull, fb);
        const unsigned act = __ballot_sync(kFull, idx < b.n);
        if (lane == 0 && m) { atomicAdd(&g_tpb_stats[0], (unsigned long long) __popc(act)); atomicAdd(&g_tpb_stats[1], (unsigned long long) __popc(m)); }
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int64_t j = base + l;
            lz4v1::lz4_decode_block(b.src + b.src_off[j], b.src_len[j], b.dst + b.dst_off[j], b.dst_cap[j], b.out_len, b.status, (uint32_t) j, lane);
            __syncwarp();
        }
    }
}

}  // namespace

extern "C" void acc_debug_tpb_stats(unsigned long long *out4)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out4, g_tpb_stats, sizeof(unsigned long long) * 4);
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_tpb_stats, z, sizeof(z));
}

int g_tpb_max_ctas = 0;   // tuning (key 2): 0 = as many CTAs as blocks need

void acc_launch_lz4_decompress_tpb(const AccBatch &b, int sm_count, cudaStream_t st)
{
    int64_t ctas = (b.n + kTpbThreads - 1) / kTpbThreads;
    const int64_t max_ctas = g_tpb_max_ctas > 0 ? g_tpb_max_ctas : (int64_t) sm_count * 16;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    lz4_decompress_tpb_kernel<<<(unsigned) ctas, kTpbThreads, 0, st>>>(b);
}
