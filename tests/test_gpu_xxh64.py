"""GPU parity: batched XXH64 vs the oracle and the reference's known answers (bit-exact)."""
import numpy as np
import pytest

import aircompressor_b200 as acb

pytestmark = pytest.mark.gpu


def test_xxh64_known_answers_and_lengths(engine, oracle):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 1 << 20, dtype=np.uint8)
    lens = list(range(0, 200)) + [222, 1023, 1024, 1025, 65535, 65536, 65537, 131072, 1 << 20]
    offs, ln = [], []
    for i, n in enumerate(lens):
        o = (i * 7919) % max(1, (1 << 20) - n + 1)   # all alignments
        offs.append(o)
        ln.append(n)
    out_len, status = engine.run_host(acb.OP_XXH64, data, np.array(offs), np.array(ln), None, None, None)
    for i, n in enumerate(lens):
        want = oracle.xxh64(data[offs[i]:offs[i] + n].tobytes(), 0)
        assert int(out_len[i]) & 0xFFFFFFFFFFFFFFFF == want, (i, n)


def test_xxh64_seeded_single_call(oracle):
    h = acb.XxHash64CudaHasher()
    PRIME32 = 0x9E3779B1
    assert h.hash(b"") == 0xEF46DB3751D8E999                      # AbstractTestXxHash64.java:65-72
    assert h.hash(b"", seed=PRIME32) == 0xAC75FDA2929B17EF
    data = bytes(range(256)) * 33
    for seed in (0, 1, PRIME32, 0xFFFFFFFFFFFFFFFF):
        assert h.hash(data, 3, 4000, seed) == oracle.xxh64(data[3:4003], seed)


def test_xxh64_hash_long(oracle):
    # AbstractTestXxHash64.java:321-364: hash(long) == hash of the 8 little-endian bytes, with and without seed
    h = acb.XxHash64CudaHasher()
    PRIME32 = 0x9E3779B1
    value = 0x0102030405060708
    assert h.hashLong(value) == h.hash(bytes([8, 7, 6, 5, 4, 3, 2, 1]))
    assert h.hashLong(value, PRIME32) == h.hash(bytes([8, 7, 6, 5, 4, 3, 2, 1]), seed=PRIME32)
    for v in (0, 12345, 0x7FFFFFFFFFFFFFFF, 0xDEADBEEFCAFEBABE):
        for seed in (0, PRIME32):
            assert h.hashLong(v, seed) == oracle.xxh64_long(v, seed)
    assert h.hashLong(0) != h.hashLong(0, PRIME32)
