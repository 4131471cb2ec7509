// zstd.cu -- placeholder for the Zstandard encode kernel (reports ACC_E_UNSUPPORTED per block until it lands).
#include "acc_device.cuh"

namespace {
__global__ void zstd_unsupported_kernel(AccBatch b)
{
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.n) { b.out_len[i] = 0; b.status[i] = ACC_STATUS(ACC_E_UNSUPPORTED, 0); }
}
}  // namespace

void acc_launch_zstd_compress(const AccBatch &b, int, cudaStream_t st, void *, int64_t)
{
    zstd_unsupported_kernel<<<(unsigned) ((b.n + 255) / 256), 256, 0, st>>>(b);
}
int64_t acc_zstd_enc_scratch_bytes(int) { return 0; }
