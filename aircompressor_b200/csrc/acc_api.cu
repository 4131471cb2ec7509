// acc_api.cu -- host runtime and C ABI of libaircompress_cuda.so (see include/aircompress_cuda.h).
//
// There is deliberately no CPU code path for any codec in this file: every compress / decompress /
// hash request becomes a kernel launch, and when no GPU is usable acc_init() fails.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include <sched.h>
#include "acc_device.cuh"

struct acc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    unsigned int *counters = nullptr;   // ring of work-stealing counters
    int counter_next = 0;
    static constexpr int kCounters = 256;
    // grow-only staging for the host-pointer entry points
    uint8_t *d_src = nullptr; int64_t d_src_cap = 0;
    uint8_t *d_dst = nullptr; int64_t d_dst_cap = 0;
    int64_t *d_idx = nullptr; int64_t d_idx_cap = 0;   // src_off | src_len | dst_off | dst_cap | out_len, then status
    void *d_scratch = nullptr; int64_t d_scratch_cap = 0;
    int64_t *h_idx = nullptr; int64_t h_idx_cap = 0;   // pinned mirror of d_idx
    uint8_t *h_stage = nullptr; int64_t h_stage_cap = 0;   // pinned staging of the single-block calls (input, then the output window)
    int32_t last_status = 0;
    int64_t last_offset = 0;
    int64_t launches = 0;
    // everything a context owns on the device (scratch, work counters, staging) is shared by its batches: work enqueued on
    // a different stream than the previous batch first waits for that batch (ev_order), so batches of one context never
    // overlap -- contexts are the unit of concurrency (one per thread, like the reference's codec objects)
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    cudaEvent_t ev_order = nullptr;
    int64_t stats[ACC_STATS_WORDS] = {};
    int tuning_ctas_per_sm = 0;
    int tuning_decoder = 0;   // LZ4 / Snappy decode: 0 = automatic (see enqueue), 1 = step decoder, 2 = record path
    int tuning_pipeline = 0;  // host-pointer batches: 0 = auto, 1 = never split, k > 1 = split into k chunks
    // copy streams + events of the pipelined host-pointer path (created on first use)
    static constexpr int kMaxChunks = 16;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_done[kMaxChunks] = {}, ev_start = nullptr;
};

static thread_local int32_t t_init_error = 0;
static const char *zstd_reason_text(int32_t reason);   // zstd_host.inc
static int64_t zstd_scratch_bytes(int32_t op, int64_t n, int sm_count);

#define CU_TRY(expr, fail_stmt) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { fail_stmt; } } while (0)

extern "C" {

int32_t acc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int32_t acc_init_error(void) { return t_init_error; }

acc_ctx *acc_init(int32_t device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        t_init_error = ACC_STATUS(ACC_E_CUDA, (int) (e != cudaSuccess ? e : cudaErrorInvalidDevice));
        return nullptr;
    }
    acc_ctx *c = new acc_ctx();
    c->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    ok = ok && cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&c->counters, sizeof(unsigned int) * acc_ctx::kCounters) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        t_init_error = ACC_STATUS(ACC_E_CUDA, (int) cudaGetLastError());
        delete c;
        return nullptr;
    }
    t_init_error = 0;
    return c;
}

void acc_destroy(acc_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    for (int i = 0; i < acc_ctx::kMaxChunks; i++) { if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]); if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]); }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_order) cudaEventDestroy(c->ev_order);
    cudaFree(c->counters); cudaFree(c->d_src); cudaFree(c->d_dst); cudaFree(c->d_idx); cudaFree(c->d_scratch);
    if (c->h_idx) cudaFreeHost(c->h_idx);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    delete c;
}

void *acc_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, (size_t) (bytes > 0 ? bytes : 1)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void acc_host_free(void *p) { if (p) cudaFreeHost(p); }

int32_t acc_last_error(acc_ctx *c, int64_t *offset)
{
    if (!c) return ACC_STATUS(ACC_E_ARGUMENT, 0);
    if (offset) *offset = c->last_offset;
    return c->last_status;
}

// NUMA node of a CUDA device from sysfs (-1: unknown / not a NUMA machine)
static int device_numa_node(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int) sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *p = bus; *p; p++) if (*p >= 'A' && *p <= 'Z') *p = (char) (*p - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int32_t acc_device_numa_node(int32_t device) { return device_numa_node(device); }

// Restricts the CALLING THREAD to the CPUs of the device's NUMA node, so that what it allocates and pins afterwards
// (first touch) and the staging copies it makes run next to the GPU's PCIe root.  Returns the node, or -1 when the
// machine gives no answer (nothing is changed then).
int32_t acc_bind_host_thread(int32_t device)
{
    const int node = device_numa_node(device);
    if (node < 0) return -1;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return -1;
    cpu_set_t want, have;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return -1;
    int any = 0;
    for (char *p = list; *p;) {                              // "0-31,64-95"
        char *e;
        long a = strtol(p, &e, 10);
        if (e == p) break;
        long b2 = a;
        if (*e == '-') { p = e + 1; b2 = strtol(p, &e, 10); }
        for (long c = a; c <= b2 && c < CPU_SETSIZE; c++) if (CPU_ISSET((int) c, &have)) { CPU_SET((int) c, &want); any++; }
        p = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    if (!any || sched_setaffinity(0, sizeof(want), &want) != 0) return -1;   // keep what the process was given if the node has none of it
    return node;
}

int32_t acc_sm_count(acc_ctx *c) { return c ? c->sm_count : 0; }
int64_t acc_kernel_launches(acc_ctx *c) { return c ? c->launches : 0; }

int32_t acc_get_stats(acc_ctx *c, int64_t *out, int32_t words)
{
    if (!c || !out || words < 0) return 0;
    c->stats[ACC_STAT_LAUNCHES] = c->launches;
    const int32_t n = words < ACC_STATS_WORDS ? words : ACC_STATS_WORDS;
    for (int32_t i = 0; i < n; i++) out[i] = c->stats[i];
    return n;
}

int32_t acc_set_tuning(acc_ctx *c, int32_t key, int32_t value)
{
    if (!c) return 0;
    if (key == 0) { int prev = c->tuning_ctas_per_sm; c->tuning_ctas_per_sm = value; return prev; }
    if (key == 1) { int prev = c->tuning_decoder; c->tuning_decoder = value; return prev; }
    if (key == 3) { int prev = c->tuning_pipeline; c->tuning_pipeline = value; return prev; }
    return 0;
}

const char *acc_code_name(int32_t code)
{
    switch (code & 0xff) {
        case ACC_OK: return "ok";
        case ACC_E_MALFORMED: return "malformed input";
        case ACC_E_DST_TOO_SMALL: return "output buffer too small";
        case ACC_E_ARGUMENT: return "illegal argument";
        case ACC_E_CUDA: return "cuda error";
        case ACC_E_UNSUPPORTED: return "unsupported";
        default: return "unknown";
    }
}

const char *acc_reason_text(int32_t reason)
{
    switch (reason) {
        case ACC_R_NONE: return "Malformed input";
        case ACC_R_INPUT_EMPTY: return "input is empty";
        case ACC_R_LAST_LITERAL_OUTSIDE: return "attempt to write last literal outside of destination buffer";
        case ACC_R_ALL_INPUT_CONSUMED: return "all input must be consumed";
        case ACC_R_OFFSET_OUTSIDE: return "offset outside destination buffer";
        case ACC_R_LAST5_LITERALS: return "last 5 bytes must be literals";
        case ACC_R_LZ4_ZERO_CAPACITY: return "zero-capacity output (reference returns -1)";
        case ACC_R_SNAPPY_TRUNCATED: return "Input is truncated";
        case ACC_R_SNAPPY_VARINT_HIGHBIT: return "last byte of compressed length int has high bit set";
        case ACC_R_SNAPPY_NEG_LENGTH: return "invalid compressed length";
        case ACC_R_SNAPPY_LEN_GT_CAP: return "Uncompressed length %s must be less than %s";
        case ACC_R_SNAPPY_LEN_MISMATCH: return "Recorded length is %s bytes but actual length after decompression is %s bytes ";
        case ACC_R_MAX_OUTPUT_TOO_SMALL: return "Max output length must be larger than maxCompressedLength";
        case ACC_R_MAX_INPUT_EXCEEDED: return "Max input length exceeded";
        default: break;
    }
    return zstd_reason_text(reason);
}

int64_t acc_lz4_compress_bound(int64_t n) { return n + n / 255 + 16; }          // Lz4RawCompressor.java:64-67
int64_t acc_snappy_compress_bound(int64_t n) { return 32 + n + n / 6; }         // SnappyRawCompressor.java:47-70
int64_t acc_zstd_compress_bound(int64_t n)                                       // ZstdJavaCompressor.java:31-40
{
    int64_t r = n + (n >> 8);
    if (n < 128 * 1024) r += (128 * 1024 - n) >> 11;
    return r;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
static unsigned int *next_counter(acc_ctx *c, cudaStream_t st)
{
    unsigned int *p = c->counters + c->counter_next;
    c->counter_next = (c->counter_next + 1) % acc_ctx::kCounters;
    cudaMemsetAsync(p, 0, sizeof(unsigned int), st);
    return p;
}


static bool grow(void **p, int64_t *cap, int64_t need, bool host)
{
    if (need <= *cap) return true;
    int64_t ncap = need + need / 4 + 4096;
    if (*p) { if (host) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    cudaError_t e = host ? cudaMallocHost(p, (size_t) ncap) : cudaMalloc(p, (size_t) ncap);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    // device staging starts defined: whole output windows are copied back to the caller (bytes behind the produced length
    // are unspecified by contract, but they should never be another allocation's leftovers)
    if (!host) cudaMemset(*p, 0, (size_t) ncap);
    *cap = ncap;
    return true;
}

// enqueue one batch kernel; all pointers in b are device pointers
static int32_t enqueue(acc_ctx *c, int32_t op, AccBatch b, cudaStream_t st, uint64_t seed)
{
    if (c->have_last && c->last_stream != st) cudaStreamWaitEvent(st, c->ev_order, 0);   // the previous batch of this context ran elsewhere
    b.work_counter = next_counter(c, st);
    switch (op) {
        case ACC_OP_LZ4_COMPRESS: acc_launch_lz4_compress(b, c->sm_count, st, next_counter(c, st)); c->launches++; break;
        case ACC_OP_LZ4_DECOMPRESS:
        case ACC_OP_SNAPPY_DECOMPRESS: {
            void *scratch = nullptr;
            unsigned int *second = nullptr;
            // The record path parses every block of the batch in one go: a fixed ~8 ms (one block's token chain, lane-serial)
            // whatever the batch size, then executes faster than the step decoder.  That pays for Snappy (whose step decoder
            // spends 6.4 instructions per byte) once the batch is large: 4 GiB of 64 KiB blocks 159 vs 130 GiB/s, 1 GiB 73 vs 115.
            const bool record_path = c->tuning_decoder == 2 || (c->tuning_decoder == 0 && op == ACC_OP_SNAPPY_DECOMPRESS && b.n >= 49152);
            if (record_path) {
                if (!grow(&c->d_scratch, &c->d_scratch_cap, acc_lz_records_scratch_bytes(b.n), false)) return -ACC_STATUS(ACC_E_CUDA, (int) cudaErrorMemoryAllocation);
                scratch = c->d_scratch;
                second = next_counter(c, st);
                c->launches++;
            }
            if (op == ACC_OP_LZ4_DECOMPRESS) acc_launch_lz4_decompress(b, c->sm_count, c->tuning_ctas_per_sm, st, scratch, second);
            else acc_launch_snappy_decompress(b, c->sm_count, c->tuning_ctas_per_sm, st, scratch, second);
            break;
        }
        case ACC_OP_SNAPPY_COMPRESS: acc_launch_snappy_compress(b, c->sm_count, st); break;
        case ACC_OP_XXH64: acc_launch_xxh64(b, seed, c->sm_count, st); break;
        case ACC_OP_XXH32: acc_launch_xxh32(b, (uint32_t) seed, c->sm_count, st); break;
        case ACC_OP_ZSTD_COMPRESS:
        case ACC_OP_ZSTD_DECOMPRESS: {
            int64_t need = zstd_scratch_bytes(op, b.n, c->sm_count);
            if (!grow(&c->d_scratch, &c->d_scratch_cap, need, false)) return -ACC_STATUS(ACC_E_CUDA, (int) cudaErrorMemoryAllocation);
            if (op == ACC_OP_ZSTD_COMPRESS) acc_launch_zstd_compress(b, c->sm_count, st, c->d_scratch, c->d_scratch_cap);
            else acc_launch_zstd_decompress(b, c->sm_count, c->tuning_ctas_per_sm, st, c->d_scratch, c->d_scratch_cap);
            break;
        }
        default: return -ACC_STATUS(ACC_E_ARGUMENT, 0);
    }
    c->launches++;
    cudaEventRecord(c->ev_order, st);
    c->last_stream = st;
    c->have_last = true;
    c->stats[ACC_STAT_BATCHES]++;
    c->stats[ACC_STAT_BLOCKS] += b.n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -ACC_STATUS(ACC_E_CUDA, (int) e);
    return 0;
}

// ---- pipelined host-pointer path ---------------------------------------------------------------------
// A large batch is cut into runs of consecutive blocks; run k's input upload (copy stream 1), its kernel (the
// caller's stream) and its output download (copy stream 2) overlap with the neighbours' so that the call costs about
// max(H2D, kernels, D2H) instead of their sum. Runs hold >= 4096 blocks: the decoders give one warp per block, and a
// launch with far fewer blocks than resident warps is bound by single-block latency, not throughput.
static int pipeline_chunks(acc_ctx *c, int64_t n, int64_t bytes, const int64_t *src_off)
{
    if (c->tuning_pipeline == 1) return 1;
    int64_t k = c->tuning_pipeline > 1 ? c->tuning_pipeline : std::min<int64_t>(n / 4096, bytes / (32 << 20));
    if (k > acc_ctx::kMaxChunks) k = acc_ctx::kMaxChunks;
    if (k > n) k = n;
    if (k < 2) return 1;
    for (int64_t i = 1; i < n; i++) if (src_off[i] < src_off[i - 1]) return 1;   // runs must cover ascending input ranges
    if (!c->s_h2d) {
        if (cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return 1; }
        for (int i = 0; i < acc_ctx::kMaxChunks; i++) {
            if (cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return 1; }
        }
    }
    return (int) k;
}

static int32_t batch_host_pipelined(acc_ctx *c, int32_t op, const void *src_base, const int64_t *src_off, const int64_t *src_len,
                                    void *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t n, cudaStream_t st,
                                    uint64_t seed, int chunks, int64_t src_lo, int64_t src_pad, int64_t dst_lo, int64_t dst_pad, bool has_dst)
{
    int64_t *h = c->h_idx;
    int32_t rc = 0;
#define PIPE_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess && rc == 0) rc = -ACC_STATUS(ACC_E_CUDA, (int) e_); } while (0)
    PIPE_TRY(cudaEventRecord(c->ev_start, st));                 // order after whatever the caller queued on st
    PIPE_TRY(cudaStreamWaitEvent(c->s_h2d, c->ev_start, 0));
    PIPE_TRY(cudaMemcpyAsync(c->d_idx, h, (size_t) (4 * n * 8), cudaMemcpyHostToDevice, c->s_h2d));
    for (int k = 0; k < chunks && rc == 0; k++) {
        const int64_t b0 = n * k / chunks, b1 = n * (k + 1) / chunks, m = b1 - b0;
        int64_t lo = src_off[b0], hi = lo;
        for (int64_t i = b0; i < b1; i++) hi = std::max(hi, src_off[i] + src_len[i]);
        if (hi > lo)
            PIPE_TRY(cudaMemcpyAsync(c->d_src + src_pad + (lo - src_lo), (const uint8_t *) src_base + lo, (size_t) (hi - lo),
                                     cudaMemcpyHostToDevice, c->s_h2d));
        PIPE_TRY(cudaEventRecord(c->ev_in[k], c->s_h2d));
        PIPE_TRY(cudaStreamWaitEvent(st, c->ev_in[k], 0));
        AccBatch b{c->d_src, c->d_idx + b0, c->d_idx + n + b0, c->d_dst, c->d_idx + 2 * n + b0, c->d_idx + 3 * n + b0,
                   c->d_idx + 4 * n + b0, (int32_t *) (c->d_idx + 5 * n) + b0, m, nullptr};
        if (rc == 0) rc = enqueue(c, op, b, st, seed);
        PIPE_TRY(cudaEventRecord(c->ev_done[k], st));
        PIPE_TRY(cudaStreamWaitEvent(c->s_d2h, c->ev_done[k], 0));
        if (has_dst && rc == 0) {
            int64_t run_lo = dst_off[b0], run_hi = dst_off[b0] + dst_cap[b0];
            for (int64_t i = b0 + 1; i <= b1; i++) {
                if (i < b1 && dst_off[i] == run_hi) { run_hi += dst_cap[i]; continue; }
                if (run_hi > run_lo)
                    PIPE_TRY(cudaMemcpyAsync((uint8_t *) dst_base + run_lo, c->d_dst + dst_pad + (run_lo - dst_lo), (size_t) (run_hi - run_lo),
                                             cudaMemcpyDeviceToHost, c->s_d2h));
                if (i < b1) { run_lo = dst_off[i]; run_hi = dst_off[i] + dst_cap[i]; }
            }
        }
    }
    if (rc == 0) PIPE_TRY(cudaMemcpyAsync(h + 4 * n, c->d_idx + 4 * n, (size_t) (n * 8 + n * 4), cudaMemcpyDeviceToHost, c->s_d2h));
    // always drain all three streams, also on failure, before the staging buffers can be reused
    PIPE_TRY(cudaStreamSynchronize(c->s_h2d));
    PIPE_TRY(cudaStreamSynchronize(st));
    PIPE_TRY(cudaStreamSynchronize(c->s_d2h));
#undef PIPE_TRY
    return rc;
}

static int32_t batch_impl(acc_ctx *c, int32_t op, const void *src_base, const int64_t *src_off, const int64_t *src_len,
                          void *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *out_len, int32_t *status,
                          int64_t n, int32_t flags, int64_t stream, uint64_t seed)
{
    if (!c || n < 0 || op < 0 || op > ACC_OP_XXH32) return -ACC_STATUS(ACC_E_ARGUMENT, 0);
    if (cudaSetDevice(c->device) != cudaSuccess) return -ACC_STATUS(ACC_E_CUDA, (int) cudaGetLastError());
    cudaStream_t st = stream ? (cudaStream_t) (uintptr_t) stream : c->stream;
    if (n == 0) return 0;

    if (flags & ACC_F_DEVICE_POINTERS) {
        AccBatch b{(const uint8_t *) src_base, src_off, src_len, (uint8_t *) dst_base, dst_off, dst_cap, out_len, status, n, nullptr};
        return enqueue(c, op, b, st, seed);
    }

    // ---- host pointers: stage, run, copy back, synchronise ----
    const bool has_dst = op != ACC_OP_XXH64 && op != ACC_OP_XXH32;
    int64_t src_lo = INT64_MAX, src_hi = 0, dst_lo = INT64_MAX, dst_hi = 0;
    for (int64_t i = 0; i < n; i++) {
        if (src_len[i] < 0 || src_off[i] < 0) return -ACC_STATUS(ACC_E_ARGUMENT, 0);
        if (src_off[i] < src_lo) src_lo = src_off[i];
        if (src_off[i] + src_len[i] > src_hi) src_hi = src_off[i] + src_len[i];
        if (has_dst) {
            if (dst_cap[i] < 0 || dst_off[i] < 0) return -ACC_STATUS(ACC_E_ARGUMENT, 0);
            if (dst_off[i] < dst_lo) dst_lo = dst_off[i];
            if (dst_off[i] + dst_cap[i] > dst_hi) dst_hi = dst_off[i] + dst_cap[i];
        }
    }
    if (src_hi < src_lo) src_hi = src_lo;
    if (!has_dst) { dst_lo = 0; dst_hi = 0; }
    if (dst_hi < dst_lo) dst_hi = dst_lo;
    const int64_t src_bytes = src_hi - src_lo, dst_bytes = dst_hi - dst_lo;
    // keep the relative alignment of the caller's buffers (mod 16) so kernels see the same layout
    const int64_t src_pad = src_lo & 15, dst_pad = dst_lo & 15;
    const int64_t idx_words = 5 * n;                       // src_off, src_len, dst_off, dst_cap, out_len
    const int64_t idx_bytes = idx_words * 8 + n * 4;       // + status
    if (!grow((void **) &c->d_src, &c->d_src_cap, src_bytes + src_pad + 64, false) ||
        !grow((void **) &c->d_dst, &c->d_dst_cap, dst_bytes + dst_pad + 64, false) ||
        !grow((void **) &c->d_idx, &c->d_idx_cap, idx_bytes, false) ||
        !grow((void **) &c->h_idx, &c->h_idx_cap, idx_bytes, true)) {
        return -ACC_STATUS(ACC_E_CUDA, (int) cudaErrorMemoryAllocation);
    }
    int64_t *h = c->h_idx;
    for (int64_t i = 0; i < n; i++) {
        h[i] = src_off[i] - src_lo + src_pad;
        h[n + i] = src_len[i];
        h[2 * n + i] = has_dst ? dst_off[i] - dst_lo + dst_pad : 0;
        h[3 * n + i] = has_dst ? dst_cap[i] : 0;
    }
    const auto t_call = std::chrono::steady_clock::now();
    auto account = [&](int64_t d2h) {
        c->stats[ACC_STAT_HOST_CALLS]++;
        c->stats[ACC_STAT_H2D_BYTES] += src_bytes + 4 * n * 8;
        c->stats[ACC_STAT_D2H_BYTES] += d2h + n * 12;
        const int64_t us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_call).count();
        c->stats[ACC_STAT_LAST_CALL_US] = us;
        c->stats[ACC_STAT_TOTAL_CALL_US] += us;
    };
    if (n == 1 && src_bytes + dst_bytes <= (8 << 20)) {
        // ---- single block (the Java-shaped entry points): pageable caller memory is staged through a pinned buffer, the
        // whole output window comes back in the same stream as the result words: ONE synchronisation per call, and only the
        // bytes produced are copied into the caller's buffer (bytes beyond the returned length stay untouched, like with
        // the reference codecs)
        if (!grow((void **) &c->h_stage, &c->h_stage_cap, src_bytes + dst_bytes + 64, true)) return -ACC_STATUS(ACC_E_CUDA, (int) cudaErrorMemoryAllocation);
        if (src_bytes > 0) memcpy(c->h_stage, (const uint8_t *) src_base + src_lo, (size_t) src_bytes);
        CU_TRY(cudaMemcpyAsync(c->d_idx, h, (size_t) (4 * 8), cudaMemcpyHostToDevice, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
        if (src_bytes > 0)
            CU_TRY(cudaMemcpyAsync(c->d_src + src_pad, c->h_stage, (size_t) src_bytes, cudaMemcpyHostToDevice, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
        AccBatch b1{c->d_src, c->d_idx, c->d_idx + 1, c->d_dst, c->d_idx + 2, c->d_idx + 3, c->d_idx + 4, (int32_t *) (c->d_idx + 5), 1, nullptr};
        const int32_t r1 = enqueue(c, op, b1, st, seed);
        if (r1 != 0) return r1;
        CU_TRY(cudaMemcpyAsync(h + 4, c->d_idx + 4, (size_t) (8 + 4), cudaMemcpyDeviceToHost, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
        if (has_dst && dst_bytes > 0)
            CU_TRY(cudaMemcpyAsync(c->h_stage + src_bytes, c->d_dst + dst_pad, (size_t) dst_bytes, cudaMemcpyDeviceToHost, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
        CU_TRY(cudaStreamSynchronize(st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
        const int64_t produced = h[4];
        const int32_t stt = ((int32_t *) (h + 5))[0];
        if (has_dst && stt == 0 && produced > 0) memcpy((uint8_t *) dst_base + dst_lo, c->h_stage + src_bytes, (size_t) (produced < dst_bytes ? produced : dst_bytes));
        out_len[0] = produced;
        if (status) status[0] = stt;
        account(dst_bytes);
        return 0;
    }
    const int chunks = pipeline_chunks(c, n, src_bytes + dst_bytes, src_off);
    if (chunks > 1) {
        int32_t r = batch_host_pipelined(c, op, src_base, src_off, src_len, dst_base, dst_off, dst_cap, n, st, seed, chunks,
                                         src_lo, src_pad, dst_lo, dst_pad, has_dst);
        if (r != 0) return r;
        memcpy(out_len, h + 4 * n, (size_t) (n * 8));
        if (status) memcpy(status, h + 5 * n, (size_t) (n * 4));
        account(dst_bytes);
        return 0;
    }
    CU_TRY(cudaMemcpyAsync(c->d_idx, h, (size_t) (4 * n * 8), cudaMemcpyHostToDevice, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
    if (src_bytes > 0)
        CU_TRY(cudaMemcpyAsync(c->d_src + src_pad, (const uint8_t *) src_base + src_lo, (size_t) src_bytes, cudaMemcpyHostToDevice, st),
               return -ACC_STATUS(ACC_E_CUDA, (int) e_));
    AccBatch b{c->d_src, c->d_idx, c->d_idx + n, c->d_dst, c->d_idx + 2 * n, c->d_idx + 3 * n, c->d_idx + 4 * n,
               (int32_t *) (c->d_idx + 5 * n), n, nullptr};
    int32_t r = enqueue(c, op, b, st, seed);
    if (r != 0) return r;
    CU_TRY(cudaMemcpyAsync(h + 4 * n, c->d_idx + 4 * n, (size_t) (n * 8 + n * 4), cudaMemcpyDeviceToHost, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
    if (has_dst && dst_bytes > 0 && n > 1) {
        // copy back every block's [dst_off, dst_off + dst_cap) window and nothing outside of them: windows that
        // touch are merged into one transfer (a gap-free batch is a single D2H copy)
        int64_t run_lo = dst_off[0], run_hi = dst_off[0] + dst_cap[0];
        for (int64_t i = 1; i <= n; i++) {
            if (i < n && dst_off[i] == run_hi) { run_hi += dst_cap[i]; continue; }
            if (run_hi > run_lo)
                CU_TRY(cudaMemcpyAsync((uint8_t *) dst_base + run_lo, c->d_dst + dst_pad + (run_lo - dst_lo), (size_t) (run_hi - run_lo),
                                       cudaMemcpyDeviceToHost, st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
            if (i < n) { run_lo = dst_off[i]; run_hi = dst_off[i] + dst_cap[i]; }
        }
    }
    CU_TRY(cudaStreamSynchronize(st), return -ACC_STATUS(ACC_E_CUDA, (int) e_));
    if (has_dst && dst_bytes > 0 && n == 1) {
        // single-block call (the Java-shaped entry points): copy back only the bytes produced, so that the caller's
        // buffer beyond the returned length stays untouched like it does with the reference codecs
        int64_t produced = h[4 * n];
        int32_t stt = ((int32_t *) (h + 5 * n))[0];
        if (stt == 0 && produced > 0) {
            CU_TRY(cudaMemcpy((uint8_t *) dst_base + dst_lo, c->d_dst + dst_pad, (size_t) produced, cudaMemcpyDeviceToHost),
                   return -ACC_STATUS(ACC_E_CUDA, (int) e_));
        }
    }
    memcpy(out_len, h + 4 * n, (size_t) (n * 8));
    if (status) memcpy(status, h + 5 * n, (size_t) (n * 4));
    account(dst_bytes);
    return 0;
}

static int64_t single_impl(acc_ctx *c, int32_t op, const void *src, int64_t src_len, void *dst, int64_t dst_cap, uint64_t seed)
{
    if (!c) return -ACC_STATUS(ACC_E_ARGUMENT, 0);
    if (src_len < 0 || dst_cap < 0) { c->last_status = ACC_STATUS(ACC_E_ARGUMENT, 0); c->last_offset = 0; return -c->last_status; }
    int64_t zero = 0, out_len = 0;
    int32_t status = 0;
    int32_t r = batch_impl(c, op, src, &zero, &src_len, dst, &zero, &dst_cap, &out_len, &status, 1, 0, 0, seed);
    if (r != 0) { c->last_status = -r; c->last_offset = 0; return r; }
    if (status != 0) { c->last_status = status; c->last_offset = out_len; return -(int64_t) status; }
    c->last_status = 0;
    c->last_offset = 0;
    return out_len;
}

extern "C" {

int32_t acc_batch(acc_ctx *c, int32_t op, const void *src_base, const int64_t *src_off, const int64_t *src_len,
                  void *dst_base, const int64_t *dst_off, const int64_t *dst_cap, int64_t *out_len, int32_t *status,
                  int64_t n, int32_t flags, int64_t stream)
{
    return batch_impl(c, op, src_base, src_off, src_len, dst_base, dst_off, dst_cap, out_len, status, n, flags, stream, 0);
}

#define ACC_BATCH_ALIAS(name, op)                                                                                          \
    int32_t name(acc_ctx *c, const void *sb, const int64_t *so, const int64_t *sl, void *db, const int64_t *d_o,          \
                 const int64_t *dc, int64_t *ol, int32_t *stt, int64_t n, int32_t flags, int64_t stream)                   \
    { return batch_impl(c, op, sb, so, sl, db, d_o, dc, ol, stt, n, flags, stream, 0); }
ACC_BATCH_ALIAS(acc_lz4_compress_batch, ACC_OP_LZ4_COMPRESS)
ACC_BATCH_ALIAS(acc_lz4_decompress_batch, ACC_OP_LZ4_DECOMPRESS)
ACC_BATCH_ALIAS(acc_snappy_compress_batch, ACC_OP_SNAPPY_COMPRESS)
ACC_BATCH_ALIAS(acc_snappy_decompress_batch, ACC_OP_SNAPPY_DECOMPRESS)
ACC_BATCH_ALIAS(acc_zstd_compress_batch, ACC_OP_ZSTD_COMPRESS)
ACC_BATCH_ALIAS(acc_zstd_decompress_batch, ACC_OP_ZSTD_DECOMPRESS)

int32_t acc_xxh64_batch(acc_ctx *c, const void *sb, const int64_t *so, const int64_t *sl, int64_t *hashes, int64_t n, int32_t flags, int64_t stream)
{
    return batch_impl(c, ACC_OP_XXH64, sb, so, sl, nullptr, nullptr, nullptr, hashes, nullptr, n, flags, stream, 0);
}

int32_t acc_xxh32_batch(acc_ctx *c, const void *sb, const int64_t *so, const int64_t *sl, int64_t *hashes, int64_t n, int32_t seed, int32_t flags, int64_t stream)
{
    return batch_impl(c, ACC_OP_XXH32, sb, so, sl, nullptr, nullptr, nullptr, hashes, nullptr, n, flags, stream, (uint64_t) (uint32_t) seed);
}

int64_t acc_lz4_compress(acc_ctx *c, const void *s, int64_t sl, void *d, int64_t dc) { return single_impl(c, ACC_OP_LZ4_COMPRESS, s, sl, d, dc, 0); }
int64_t acc_lz4_decompress(acc_ctx *c, const void *s, int64_t sl, void *d, int64_t dc) { return single_impl(c, ACC_OP_LZ4_DECOMPRESS, s, sl, d, dc, 0); }
int64_t acc_snappy_compress(acc_ctx *c, const void *s, int64_t sl, void *d, int64_t dc) { return single_impl(c, ACC_OP_SNAPPY_COMPRESS, s, sl, d, dc, 0); }
int64_t acc_snappy_decompress(acc_ctx *c, const void *s, int64_t sl, void *d, int64_t dc) { return single_impl(c, ACC_OP_SNAPPY_DECOMPRESS, s, sl, d, dc, 0); }
int64_t acc_zstd_compress(acc_ctx *c, const void *s, int64_t sl, void *d, int64_t dc) { return single_impl(c, ACC_OP_ZSTD_COMPRESS, s, sl, d, dc, 0); }
int64_t acc_zstd_decompress(acc_ctx *c, const void *s, int64_t sl, void *d, int64_t dc) { return single_impl(c, ACC_OP_ZSTD_DECOMPRESS, s, sl, d, dc, 0); }

int64_t acc_xxh64(acc_ctx *c, const void *src, int64_t len, int64_t seed)
{
    if (!c || len < 0) return 0;
    int64_t zero = 0, h = 0;
    int32_t r = batch_impl(c, ACC_OP_XXH64, src, &zero, &len, nullptr, nullptr, nullptr, &h, nullptr, 1, 0, 0, (uint64_t) seed);
    c->last_status = r ? -r : 0;
    return h;
}

int32_t acc_xxh32(acc_ctx *c, const void *src, int64_t len, int32_t seed)
{
    if (!c || len < 0) return 0;
    int64_t zero = 0, h = 0;
    int32_t r = batch_impl(c, ACC_OP_XXH32, src, &zero, &len, nullptr, nullptr, nullptr, &h, nullptr, 1, 0, 0, (uint64_t) (uint32_t) seed);
    c->last_status = r ? -r : 0;
    return (int32_t) (uint32_t) h;
}

// SnappyRawDecompressor.readUncompressedLength (snappy/SnappyRawDecompressor.java:277-321): header-only, host side.
int64_t acc_snappy_uncompressed_length(const void *src, int64_t src_len, int64_t *err_offset)
{
    const uint8_t *in = (const uint8_t *) src;
    uint32_t result = 0;
    int64_t n = 0;
    for (int shift = 0;; shift += 7) {
        if (n >= src_len) { if (err_offset) *err_offset = src_len - n; return -(int64_t) ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_TRUNCATED); }
        uint32_t b = in[n++];
        result |= (b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        if (shift == 28) { if (err_offset) *err_offset = n; return -(int64_t) ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_VARINT_HIGHBIT); }
    }
    if ((int32_t) result < 0) { if (err_offset) *err_offset = 0; return -(int64_t) ACC_STATUS(ACC_E_MALFORMED, ACC_R_SNAPPY_NEG_LENGTH); }
    return (int64_t) result;
}

}  // extern "C"

#include "zstd_host.inc"
