"""Corpus access shared by bench.py and tests (not part of the product package).

Silesia-shaped inputs: the full corpus when a copy is present under corpus/silesia/ (git-ignored,
travels with gpurun), otherwise the committed stratified sample tests/golden/silesia_sample.bin
(see tests/golden/make_silesia_sample.py).  /root/reference is never read at run time.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
FILES = ["dickens", "mozilla", "mr", "nci", "ooffice", "osdb", "reymont", "samba", "sao", "webster", "x-ray", "xml"]


def load_pieces():
    """Returns (label, [numpy uint8 arrays]) -- the files (full corpus) or the sampled 128 KiB pieces."""
    full = os.path.join(ROOT, "corpus", "silesia")
    if all(os.path.exists(os.path.join(full, f)) for f in FILES):
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
