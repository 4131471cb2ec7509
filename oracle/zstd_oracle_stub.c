#include "oracle.h"
int64_t orc_zstd_max_compressed_length(int64_t n) { return -1; }
int64_t orc_zstd_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap) { return -255; }
int64_t orc_zstd_decompress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, int64_t *e) { return -255; }
int64_t orc_zstd_decompressed_size(const uint8_t *in, int64_t in_len, int64_t *e) { return -255; }
