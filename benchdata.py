"""Corpus access shared by bench.py, __graft_entry__ and tests (not part of the product package).

The workload of BASELINE.md s3 is the Silesia corpus (12 files, 211,938,580 bytes; 3,222 blocks of 64 KiB).  The files
are committed xz-compressed under data/silesia_xz/ (48 MB; SHA256SUMS beside them) and unpacked on first use into
corpus/silesia/ (git-ignored and gpurun-ignored: every box unpacks its own copy, ~3 s).  /root/reference is never read
at run time.  The 1/24 stratified sample of round 1 (tests/golden/silesia_sample.bin) remains for the small CPU tests.
"""
import hashlib
import json
import lzma
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
FILES = ["dickens", "mozilla", "mr", "nci", "ooffice", "osdb", "reymont", "samba", "sao", "webster", "x-ray", "xml"]
XZ_DIR = os.path.join(ROOT, "data", "silesia_xz")
FULL_DIR = os.path.join(ROOT, "corpus", "silesia")


def _unpack_one(name):
    dst = os.path.join(FULL_DIR, name)
    if os.path.exists(dst):
        return
    with lzma.open(os.path.join(XZ_DIR, name + ".xz"), "rb") as f:
        data = f.read()
    tmp = f"{dst}.{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, dst)   # atomic: concurrent ranks may unpack at the same time


def ensure_corpus(verify=False):
    """Unpacks data/silesia_xz/*.xz into corpus/silesia/ when files are missing.  Returns the directory, or None when the
    compressed corpus is not in the tree either."""
    if not all(os.path.exists(os.path.join(FULL_DIR, f)) for f in FILES):
        if not all(os.path.exists(os.path.join(XZ_DIR, f + ".xz")) for f in FILES):
            return None
        os.makedirs(FULL_DIR, exist_ok=True)
        with ThreadPoolExecutor(max_workers=6) as ex:   # lzma releases the GIL
            list(ex.map(_unpack_one, FILES))
    if verify:
        sums = dict(line.split()[::-1] for line in open(os.path.join(XZ_DIR, "SHA256SUMS")))
        for f in FILES:
            h = hashlib.sha256(open(os.path.join(FULL_DIR, f), "rb").read()).hexdigest()
            assert h == sums[f], f"corpus file {f} does not match data/silesia_xz/SHA256SUMS"
    return FULL_DIR


def load_pieces(sample=False):
    """Returns (label, [numpy uint8 arrays]): the 12 corpus files, or (sample=True, or no corpus in the tree) the 1/24
    stratified sample of 128 KiB pieces."""
    full = None if sample else ensure_corpus()
    if full:
        return "silesia (full corpus, 12 files)", [np.fromfile(os.path.join(full, f), dtype=np.uint8) for f in FILES]
    blob = np.fromfile(os.path.join(ROOT, "tests", "golden", "silesia_sample.bin"), dtype=np.uint8)
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "silesia_sample.json")))
    pieces = [blob[p["at"]:p["at"] + p["length"]] for p in meta["pieces"]]
    return f"silesia stratified sample ({len(pieces)} x 128 KiB pieces, every 24th chunk of the corpus)", pieces


def cut_blocks(pieces, block_size):
    """Cuts every piece into fixed-size blocks in order, keeping the final short block.
    Pieces shorter than block_size are first concatenated so large block sizes still get full blocks."""
    if block_size > 128 * 1024 and all(p.size <= 128 * 1024 for p in pieces):
        pieces = [np.concatenate(pieces)]
    out = []
    for p in pieces:
        for off in range(0, p.size, block_size):
            out.append(p[off:off + block_size])
    return out


def pack(blocks):
    """list of arrays -> (packed uint8 array, offsets int64, lengths int64)"""
    lens = np.array([b.size for b in blocks], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64) if len(blocks) else np.zeros(0, np.int64)
    packed = np.concatenate(blocks) if blocks else np.zeros(0, np.uint8)
    return packed, offs, lens
