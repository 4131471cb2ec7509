#include "oracle.h"
int64_t orc_zstd_max_compressed_length(int64_t n) { int64_t r = n + (n >> 8); if (n < 131072) r += (131072 - n) >> 11; return r; }
int64_t orc_zstd_compress(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap) { return -255; }
