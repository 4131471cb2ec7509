/*
 * examples/acc_roundtrip.c -- a plain-C consumer of the C ABI (include/aircompress_cuda.h): compresses a file in 64 KiB
 * blocks as ONE batch, decompresses it again and compares.  This is the call sequence a JNI/FFM/cgo host makes.
 *
 *   gcc -O2 -I include examples/acc_roundtrip.c -L aircompressor_b200 -laircompress_cuda -Wl,-rpath,'$ORIGIN/../aircompressor_b200' -o acc_roundtrip
 *   ./acc_roundtrip <file> [lz4|snappy|zstd]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aircompress_cuda.h"

static void die(const char *what) { fprintf(stderr, "%s\n", what); exit(1); }

int main(int argc, char **argv)
{
    if (argc < 2) die("usage: acc_roundtrip <file> [lz4|snappy|zstd]");
    const char *codec = argc > 2 ? argv[2] : "lz4";
    const int cop = !strcmp(codec, "snappy") ? ACC_OP_SNAPPY_COMPRESS : !strcmp(codec, "zstd") ? ACC_OP_ZSTD_COMPRESS : ACC_OP_LZ4_COMPRESS;
    const int dop = cop + 1;
    const int64_t block = !strcmp(codec, "zstd") ? 128 * 1024 : 64 * 1024;

    FILE *f = fopen(argv[1], "rb");
    if (!f) die("cannot open input");
    fseek(f, 0, SEEK_END);
    const int64_t size = ftell(f);
    fseek(f, 0, SEEK_SET);

    acc_ctx *ctx = acc_init(0);
    if (!ctx) { fprintf(stderr, "acc_init failed: %s\n", acc_code_name(acc_init_error())); return 2; }

    /* pinned staging makes the copies run at PCIe speed and lets large batches overlap upload, kernels and download */
    uint8_t *raw = (uint8_t *) acc_host_alloc(size > 0 ? size : 1);
    if (!raw || fread(raw, 1, (size_t) size, f) != (size_t) size) die("cannot read input");
    fclose(f);

    const int64_t n = (size + block - 1) / block;
    const int64_t bound = cop == ACC_OP_LZ4_COMPRESS ? acc_lz4_compress_bound(block)
                        : cop == ACC_OP_SNAPPY_COMPRESS ? acc_snappy_compress_bound(block) : acc_zstd_compress_bound(block);
    int64_t *idx = (int64_t *) calloc((size_t) (6 * (n + 1)), sizeof(int64_t));
    int32_t *status = (int32_t *) calloc((size_t) (n + 1), sizeof(int32_t));
    int64_t *raw_off = idx, *raw_len = idx + n, *comp_off = idx + 2 * n, *comp_cap = idx + 3 * n, *comp_len = idx + 4 * n, *out_len = idx + 5 * n;
    for (int64_t i = 0; i < n; i++) {
        raw_off[i] = i * block;
        raw_len[i] = size - i * block < block ? size - i * block : block;
        comp_off[i] = i * bound;
        comp_cap[i] = bound;
    }
    uint8_t *comp = (uint8_t *) acc_host_alloc(n * bound + 1);
    uint8_t *back = (uint8_t *) acc_host_alloc(size > 0 ? size : 1);
    if (!comp || !back) die("acc_host_alloc failed");

    int32_t rc = acc_batch(ctx, cop, raw, raw_off, raw_len, comp, comp_off, comp_cap, comp_len, status, n, 0, 0);
    if (rc != 0) { fprintf(stderr, "compress batch failed: %s\n", acc_code_name(-rc)); return 3; }
    int64_t total = 0;
    for (int64_t i = 0; i < n; i++) {
        if (status[i] != 0) { fprintf(stderr, "block %lld: %s\n", (long long) i, acc_reason_text(status[i] >> 8)); return 3; }
        total += comp_len[i];
    }
    rc = acc_batch(ctx, dop, comp, comp_off, comp_len, back, raw_off, raw_len, out_len, status, n, 0, 0);
    if (rc != 0) { fprintf(stderr, "decompress batch failed: %s\n", acc_code_name(-rc)); return 4; }
    for (int64_t i = 0; i < n; i++) {
        if (status[i] != 0 || out_len[i] != raw_len[i]) {
            fprintf(stderr, "block %lld: %s at offset %lld\n", (long long) i, acc_reason_text(status[i] >> 8), (long long) out_len[i]);
            return 4;
        }
    }
    const int same = memcmp(raw, back, (size_t) size) == 0;
    printf("%s: %lld bytes in %lld blocks -> %lld bytes (ratio %.3f), round trip %s, %lld kernel launches\n", codec, (long long) size,
           (long long) n, (long long) total, size ? (double) total / (double) size : 0.0, same ? "ok" : "MISMATCH",
           (long long) acc_kernel_launches(ctx));
    acc_host_free(raw); acc_host_free(comp); acc_host_free(back);
    free(idx); free(status);
    acc_destroy(ctx);
    return same ? 0 : 5;
}
