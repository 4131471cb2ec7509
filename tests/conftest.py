import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refnative():
    from oracle.pyoracle import RefNative
    return RefNative()


@pytest.fixture(scope="session")
def pieces():
    """the committed 1/24 stratified sample (small, for the quick parity cases)"""
    import benchdata
    return benchdata.load_pieces(sample=True)[1]


@pytest.fixture(scope="session")
def corpus_files():
    """all 12 files of the Silesia corpus (unpacked on first use from data/silesia_xz/)"""
    import benchdata
    label, files = benchdata.load_pieces()
    assert "full corpus" in label, "data/silesia_xz is missing from the tree"
    return files


@pytest.fixture(scope="session")
def sample_blocks(pieces):
    """A deterministic mix of block sizes cut from the Silesia sample (bytes objects)."""
    import benchdata
    out = []
    for size, stride in ((64 * 1024, 5), (4 * 1024, 97), (128 * 1024, 7), (1000, 411)):
        blocks = benchdata.cut_blocks(pieces, size)
        out += [blocks[i].tobytes() for i in range(0, len(blocks), stride)]
    return out


SYNTHETIC = [
    b"",
    b"hello world!",
    b"XXXXabcdabcdABCDABCDwxyzwzyz123",
    b"XXXXabcdefgh abcdefgh abcdefgh abcdefgh abcdefgh abcdefgh ABC",
    bytes(range(256)),
]  # AbstractTestCompression.java:47-56


@pytest.fixture(scope="session")
def synthetic_cases():
    rng = np.random.default_rng(1234)
    extra = [
        bytes(1), b"a" * 13, b"a" * 12, b"ab" * 40, b"a" * 70000, bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
        bytes(rng.integers(0, 4, 5000, dtype=np.uint8)), (b"0123456789abcdef" * 5000)[:70001],
    ]
    return SYNTHETIC + extra


@pytest.fixture(scope="session")
def engine():
    import aircompressor_b200 as acb
    return acb.BatchEngine(0)
