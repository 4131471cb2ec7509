"""CPU restatement of the reference's Hadoop block streams for LZ4 and Snappy -- TEST INFRASTRUCTURE (only tests/ may import it).

Sequential, one chunk at a time, like the Java: the writer is Lz4HadoopOutputStream.write / finish / writeNextChunk
(lz4/Lz4HadoopOutputStream.java:58-117; the Snappy twin differs only in compressionOverhead, :128-131 of each file), the
reader is Lz4HadoopInputStream.read / bufferCompressedData / readBigEndianInt (lz4/Lz4HadoopInputStream.java:59-162), over
the oracle block codecs.  What the batched GPU streams (aircompressor_b200/hadoop_streams.py) must agree with: the framing
byte for byte (chunk boundaries, length words), the decoded bytes, and which error ends a damaged stream after which bytes.
"""
import struct


def overhead(codec, size):
    return max(int(size * 0.01), 10) if codec == "lz4" else size // 6 + 32


def write_stream(oracle, codec, data, buffer_size=256 * 1024):
    """everything written, then close(): chunks of inputMaxSize bytes, the rest at finish()"""
    chunk = buffer_size - overhead(codec, buffer_size)
    out = bytearray()
    for pos in range(0, len(data), chunk):
        piece = data[pos:pos + chunk]
        c = oracle.compress(codec, piece)
        out += struct.pack(">II", len(piece), len(c)) + c
    return bytes(out)


def chunk_lengths(stream):
    """[(uncompressed, compressed)] of a stream written with one chunk per block"""
    res, pos = [], 0
    while pos < len(stream):
        u, c = struct.unpack_from(">II", stream, pos)
        res.append((u, c))
        pos += 8 + c
    return res


def read_stream(oracle, codec, stream, buffer_size=256 * 1024):
    """-> (decoded bytes, None) or (bytes delivered before the failure, exception): the read loop until it returns -1 or throws"""
    out = bytearray()
    pos = 0
    block_left = 0

    def read_int():
        nonlocal pos
        if pos >= len(stream):
            return -1
        if pos + 4 > len(stream):
            pos = len(stream)
            raise IOError("Stream is truncated")
        v = struct.unpack_from(">I", stream, pos)[0]
        pos += 4
        return v

    try:
        while True:
            while block_left == 0:                                   # bufferCompressedData :104-110
                block_left = read_int()
                if block_left == -1:
                    return bytes(out), None
            n = read_int()
            if n == -1:
                return bytes(out), None
            if pos + n > len(stream):
                raise EOFError("encountered EOF while reading block data")
            chunk = stream[pos:pos + n]
            pos += n
            # the Java decodes into uncompressedChunk (buffer_size + 8 bytes) or straight into a user buffer that is at least
            # as long as the block; the capacity that matters for well-formed streams is the block's remaining length
            r, off, data = oracle.decompress_raw(codec, chunk, block_left)
            if r < 0:
                return bytes(out), ("block", -r, off)
            out += data[:r].tobytes()
            block_left -= r
    except (IOError, EOFError) as e:
        return bytes(out), e
