"""Generates tests/golden/silesia_sample.bin: a stratified sample of the reference's Silesia corpus
(/root/reference/testdata/silesia, the corpus of T/benchmark/DataSet.java:28-89).

Every file is cut into 128 KiB chunks in file order (the final short chunk of a file is kept, as in
BASELINE.md section 3); chunk k of the whole corpus is sampled when k % 24 == 7.  The result
(~8.6 MB, all 12 files represented proportionally to their size) travels with the repo so that GPU
tests and bench.py have Silesia-shaped data on the GPU box, where /root/reference does not exist.
Run in the build container only:  python tests/golden/make_silesia_sample.py
"""
import json
import os

SRC = "/root/reference/testdata/silesia"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["dickens", "mozilla", "mr", "nci", "ooffice", "osdb", "reymont", "samba", "sao", "webster", "x-ray", "xml"]
CHUNK = 128 * 1024


def main():
    out = bytearray()
    index = []
    k = 0
    for name in FILES:
        data = open(os.path.join(SRC, name), "rb").read()
        for off in range(0, len(data), CHUNK):
            if k % 24 == 7:
                piece = data[off:off + CHUNK]
                index.append({"file": name, "offset": off, "length": len(piece), "at": len(out)})
                out += piece
            k += 1
    open(os.path.join(HERE, "silesia_sample.bin"), "wb").write(out)
    json.dump({"chunk": CHUNK, "stride": 24, "phase": 7, "total_chunks": k, "pieces": index},
              open(os.path.join(HERE, "silesia_sample.json"), "w"), indent=0)
    print(len(index), "pieces", len(out), "bytes")


if __name__ == "__main__":
    main()
