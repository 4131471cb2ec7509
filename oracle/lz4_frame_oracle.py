"""CPU restatement of the reference's LZ4 frame codec -- TEST INFRASTRUCTURE (only tests/ may import it).

Follows lz4/Lz4FrameCompression.java statement by statement: compress :83-133, decompress :135-180, decompressFrame
:187-315, skipFrame :320-335, with the block codec = the oracle port of Lz4RawCompressor / Lz4RawDecompressor
(oracle/lz4_oracle.c) and the checksum = oracle/xxh32_oracle.c (both pinned against the reference's vectors and bundled
natives).  Sequential, one block at a time, exactly like the Java loop: this is what the batched GPU frame codec
(aircompressor_b200/lz4_frame.py) must agree with -- bytes, lengths, and for malformed input the message and the offset.
"""
import struct

MAGIC = 0x184D2204
SKIPPABLE_MAGIC, SKIPPABLE_MAGIC_MASK = 0x184D2A50, 0xFFFFFFF0
FLG_VERSION, FLG_BLOCK_INDEPENDENCE, FLG_BLOCK_CHECKSUM, FLG_CONTENT_SIZE, FLG_CONTENT_CHECKSUM, FLG_DICTIONARY_ID = 0x40, 0x20, 0x10, 0x08, 0x04, 0x01
FLG_RESERVED_MASK, BD_RESERVED_MASK = 0x02, 0x8F
BD_4MB, BLOCK_MAX_SIZE_4MB = 0x70, 4 << 20
HEADER_SIZE = 7
UNCOMPRESSED_BLOCK_FLAG, BLOCK_SIZE_MASK = 0x80000000, 0x7FFFFFFF


class FrameError(Exception):
    """MalformedInputException(offset, message) of the Java."""

    def __init__(self, offset, reason):
        super().__init__(f"{reason}: offset={offset}")
        self.offset, self.reason = offset, reason


def block_maximum_size(i):
    return {4: 64 << 10, 5: 256 << 10, 6: 1 << 20, 7: 4 << 20}.get(i, -1)      # Lz4FrameFormat.java:58-67


def compress(oracle, data):
    """Lz4FrameCompression.compress :83-133 (output buffer assumed large enough)."""
    out = bytearray(struct.pack("<I", MAGIC))
    out += bytes([FLG_VERSION | FLG_BLOCK_INDEPENDENCE, BD_4MB])
    out.append((oracle.xxh32(bytes(out[4:6])) >> 8) & 0xFF)
    pos = 0
    while pos < len(data):
        block = data[pos:pos + BLOCK_MAX_SIZE_4MB]
        c = oracle.compress("lz4", block)
        if len(c) < len(block):
            out += struct.pack("<I", len(c)) + c
        else:
            out += struct.pack("<I", len(block) | UNCOMPRESSED_BLOCK_FLAG) + block
        pos += len(block)
    out += struct.pack("<I", 0)
    return bytes(out)


def decompress(oracle, inp, max_output):
    """Lz4FrameCompression.decompress :135-180.  Returns the decoded bytes or raises FrameError."""
    n = len(inp)
    if n < HEADER_SIZE:
        raise FrameError(0, "Input is too short to be an LZ4 frame")
    out = bytearray()
    pos = 0
    while pos < n:
        if pos + 4 > n:
            raise FrameError(pos, "Truncated LZ4 frame: incomplete magic number")
        magic = struct.unpack_from("<I", inp, pos)[0]
        if magic == MAGIC:
            pos = _frame(oracle, inp, pos, out, max_output)
        elif (magic & SKIPPABLE_MAGIC_MASK) == SKIPPABLE_MAGIC:
            sp = pos + 4                                                            # skipFrame :320-335
            if sp + 4 > n:
                raise FrameError(sp, "Truncated LZ4 skippable frame: missing frame size")
            end = sp + 4 + struct.unpack_from("<I", inp, sp)[0]
            if end > n:
                raise FrameError(sp, "Truncated LZ4 skippable frame")
            pos = end
        else:
            raise FrameError(pos, "Invalid LZ4 frame magic number")
    return bytes(out)


def _frame(oracle, inp, start, out, limit):
    """decompressFrame :187-315; appends to `out`, returns the input position behind the frame."""
    n = len(inp)
    out_start = len(out)
    d0 = start + 4
    if d0 + 2 > n:
        raise FrameError(d0, "Truncated LZ4 frame header")
    flg, bd = inp[d0], inp[d0 + 1]
    version = (flg >> 6) & 3
    if version != 1:
        raise FrameError(d0, f"Unsupported LZ4 frame version: {version}")
    if (flg & FLG_RESERVED_MASK) or (bd & BD_RESERVED_MASK):
        raise FrameError(d0, "Corrupt LZ4 frame: reserved bits in the frame descriptor must be zero")
    if not flg & FLG_BLOCK_INDEPENDENCE:
        raise FrameError(d0, "LZ4 frames with linked blocks are not supported")
    if flg & FLG_DICTIONARY_ID:
        raise FrameError(d0, "LZ4 frames with a dictionary are not supported")
    bmax = block_maximum_size((bd >> 4) & 7)
    if bmax < 0:
        raise FrameError(d0 + 1, "Invalid LZ4 frame block maximum size")
    pos = d0 + 2
    has_size = bool(flg & FLG_CONTENT_SIZE)
    if pos + (8 if has_size else 0) + 1 > n:
        raise FrameError(pos, "Truncated LZ4 frame header")
    expected_size = -1
    if has_size:
        expected_size = struct.unpack_from("<q", inp, pos)[0]
        pos += 8
    if inp[pos] != (oracle.xxh32(bytes(inp[d0:pos])) >> 8) & 0xFF:
        raise FrameError(pos, "Corrupt LZ4 frame: invalid header checksum")
    pos += 1
    while True:
        if pos + 4 > n:
            raise FrameError(pos, "Truncated LZ4 frame: missing block size")
        hdr = struct.unpack_from("<I", inp, pos)[0]
        pos += 4
        if hdr == 0:
            break
        stored, length = bool(hdr & UNCOMPRESSED_BLOCK_FLAG), hdr & BLOCK_SIZE_MASK
        if length > bmax or pos + length > n:
            raise FrameError(pos, "Truncated LZ4 frame: block extends past end of input")
        if stored:
            if len(out) + length > limit:
                raise FrameError(len(out), "Output buffer too small")
            out += inp[pos:pos + length]
        else:
            # the Java hands the block decoder the rest of the output buffer; the oracle block decoder gets the same capacity
            cap = limit - len(out)
            r, off, data = oracle.decompress_raw("lz4", bytes(inp[pos:pos + length]), cap)
            if r < 0:
                raise BlockError(-r, off, len(out))
            if r > bmax:
                raise FrameError(pos, "Corrupt LZ4 frame: decompressed block exceeds maximum block size")
            out += data[:r].tobytes()
        if flg & FLG_BLOCK_CHECKSUM:
            cp = pos + length
            if cp + 4 > n:
                raise FrameError(cp, "Truncated LZ4 frame: missing block checksum")
            if struct.unpack_from("<I", inp, cp)[0] != oracle.xxh32(bytes(inp[pos:pos + length])):
                raise FrameError(cp, "Corrupt LZ4 frame: invalid block checksum")
        pos += length + (4 if flg & FLG_BLOCK_CHECKSUM else 0)
    content = len(out) - out_start
    if flg & FLG_CONTENT_CHECKSUM:
        if pos + 4 > n:
            raise FrameError(pos, "Truncated LZ4 frame: missing content checksum")
        if struct.unpack_from("<I", inp, pos)[0] != oracle.xxh32(bytes(out[out_start:])):
            raise FrameError(pos, "Corrupt LZ4 frame: invalid content checksum")
        pos += 4
    if has_size and content != expected_size:
        raise FrameError(pos, "Corrupt LZ4 frame: content size does not match frame header")
    return pos


class BlockError(Exception):
    """The raw block decoder rejected a block (its own MalformedInputException in the Java): status word, offset inside the
    block, output position of the block."""

    def __init__(self, status, offset, out_pos):
        super().__init__(f"block decoder status {status:#x} at {offset}")
        self.status, self.offset, self.out_pos = status, offset, out_pos
