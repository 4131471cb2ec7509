"""LZ4 frame format over the batched GPU block codec.

Mirror of the reference's Lz4FrameCompression (lz4/Lz4FrameCompression.java:83-371) with the class names its GPU analogue
would carry next to Lz4FrameJavaCompressor / Lz4FrameNativeCompressor (lz4/Lz4FrameJavaCompressor.java:25-44): the framing
(magic number, frame descriptor, XXH32 header checksum, block framing, end mark, optional block / content checksums and
content size, concatenated and skippable frames) is host-side bookkeeping exactly as in the reference, while

  * every data block of a call goes to the GPU as ONE batch (acc_batch, LZ4 raw block op): a frame is a list of independent
    blocks of at most 4 MiB, i.e. the natural producer of a batch from a single user call;
  * block checksums are one XXH32 batch (acc_xxh32_batch), header and content checksums one-shot XXH32 calls (acc_xxh32).

Error texts and offsets are the reference's (Lz4FrameCompression.java:140-346); because blocks are decoded as a batch, the
header walk runs first and errors are then reported in the order the Java's sequential loop would meet them.
There is no CPU fallback: without the CUDA library the constructors fail.
"""
import struct

import numpy as np

from . import _native as N
from .api import BatchEngine, IllegalArgumentException, MalformedInputException, _as_array, _Context, _verify_range

# Lz4FrameFormat.java:24-52
MAGIC = 0x184D2204
SKIPPABLE_MAGIC = 0x184D2A50
SKIPPABLE_MAGIC_MASK = 0xFFFFFFF0
FLG_VERSION = 0b01 << 6
FLG_BLOCK_INDEPENDENCE = 1 << 5
FLG_BLOCK_CHECKSUM = 1 << 4
FLG_CONTENT_SIZE = 1 << 3
FLG_CONTENT_CHECKSUM = 1 << 2
FLG_DICTIONARY_ID = 1
FLG_RESERVED_MASK = 0b00000010
BD_RESERVED_MASK = 0b10001111
BD_4MB = 7 << 4
BLOCK_MAX_SIZE_4MB = 4 * 1024 * 1024
HEADER_SIZE = 7
END_MARK_SIZE = 4
UNCOMPRESSED_BLOCK_FLAG = 0x80000000
BLOCK_SIZE_MASK = 0x7FFFFFFF
INT_MAX = 0x7FFFFFFF


def block_maximum_size(block_size_id):
    """Lz4FrameFormat.blockMaximumSize (Lz4FrameFormat.java:58-67)."""
    return {4: 64 * 1024, 5: 256 * 1024, 6: 1024 * 1024, 7: 4 * 1024 * 1024}.get(block_size_id, -1)


class _Xxh32:
    """XxHash32Hasher.hash(...) (xxhash/XxHash32Hasher.java:18-50) on the GPU: one-shot and batched."""

    def __init__(self, ctx):
        self._ctx = ctx
        self._L = N.lib()

    def hash(self, arr, offset, length, seed=0):
        return self._L.acc_xxh32(self._ctx.handle, arr.ctypes.data + offset, length, seed) & 0xFFFFFFFF

    def hash_many(self, arr, offsets, lengths, seed=0):
        n = len(offsets)
        out = np.zeros(n, dtype=np.int64)
        if n == 0:
            return out
        so, sl = np.ascontiguousarray(offsets, dtype=np.int64), np.ascontiguousarray(lengths, dtype=np.int64)
        r = self._L.acc_xxh32_batch(self._ctx.handle, arr.ctypes.data, so.ctypes.data, sl.ctypes.data, out.ctypes.data, n, seed, 0, 0)
        if r != 0:
            raise RuntimeError(f"acc_xxh32_batch failed: status {-r:#x}")
        return out & 0xFFFFFFFF


class XxHash32CudaHasher:
    """One-shot XXH32 (XxHash32Hasher.hash overloads, xxhash/XxHash32Hasher.java:18-50; XxHash32JavaHasher.java:68-109)."""

    def __init__(self, device=0):
        self._ctx = _Context(device)
        self._x = _Xxh32(self._ctx)

    def hash(self, input, offset=0, length=None, seed=0):
        src = _as_array(input)
        if length is None:
            length = src.size - offset
        _verify_range(src, offset, length)
        s = seed & 0xFFFFFFFF
        h = self._x.hash(src, offset, length, s - (1 << 32) if s >= (1 << 31) else s)
        return h - (1 << 32) if h >= (1 << 31) else h          # a Java int

    def hash_many(self, input, offsets, lengths, seed=0):
        return self._x.hash_many(_as_array(input), offsets, lengths, seed)

    def close(self):
        self._ctx.close()


class Lz4FrameCudaCompressor:
    """Lz4FrameCompressor (lz4/Lz4FrameCompressor.java:21-31) whose blocks are compressed as one GPU batch."""

    def __init__(self, device=0):
        self._engine = BatchEngine(device)
        self._L = N.lib()

    def maxCompressedLength(self, uncompressedSize):
        """Lz4FrameCompression.maxCompressedLength (Lz4FrameCompression.java:67-80)."""
        if uncompressedSize < 0:
            raise IllegalArgumentException(f"uncompressedSize is negative: {uncompressedSize}")
        blocks = (uncompressedSize + BLOCK_MAX_SIZE_4MB - 1) // BLOCK_MAX_SIZE_4MB
        max_length = HEADER_SIZE + END_MARK_SIZE + uncompressedSize + 4 * blocks
        if max_length > INT_MAX:
            raise IllegalArgumentException(f"Maximum compressed length exceeds Integer.MAX_VALUE for uncompressedSize: {uncompressedSize}")
        return max_length

    def compress(self, input, inputOffset, inputLength, output, outputOffset, maxOutputLength):
        """Lz4FrameCompression.compress (Lz4FrameCompression.java:83-133): one frame, independent blocks of at most 4 MiB, no
        checksums or content size -- what the reference writes."""
        src = _as_array(input)
        dst = _as_array(output, writable=True)
        _verify_range(src, inputOffset, inputLength)
        _verify_range(dst, outputOffset, maxOutputLength)
        out = dst[outputOffset:outputOffset + maxOutputLength]
        pos = 0

        def need(n):
            if pos + n > maxOutputLength:
                raise IllegalArgumentException("Output buffer too small")          # ensureCapacity :355-360

        # frame header :96-101 (the header checksum covers FLG and BD)
        need(4); out[pos:pos + 4] = np.frombuffer(struct.pack("<I", MAGIC), dtype=np.uint8); pos += 4
        need(1); out[pos] = FLG_VERSION | FLG_BLOCK_INDEPENDENCE; pos += 1
        need(1); out[pos] = BD_4MB; pos += 1
        hc = (self._L.acc_xxh32(self._engine._ctx.handle, out.ctypes.data + pos - 2, 2, 0) >> 8) & 0xFF
        need(1); out[pos] = hc; pos += 1

        # data blocks :104-127 -- all of them in one batch
        n = (inputLength + BLOCK_MAX_SIZE_4MB - 1) // BLOCK_MAX_SIZE_4MB
        if n:
            so = inputOffset + np.arange(n, dtype=np.int64) * BLOCK_MAX_SIZE_4MB
            sl = np.minimum(BLOCK_MAX_SIZE_4MB, inputOffset + inputLength - so)
            bound = int(self._L.acc_lz4_compress_bound(int(sl.max())))
            scratch = np.empty(n * bound, dtype=np.uint8)
            do = np.arange(n, dtype=np.int64) * bound
            clen, st = self._engine.run_host(N.OP_LZ4_COMPRESS, src, so, sl, scratch, do, np.full(n, bound, dtype=np.int64))
            if (st != 0).any():
                raise RuntimeError(f"LZ4 block compression failed: status {int(st[st != 0][0]):#x}")
            for i in range(n):
                bl, cl = int(sl[i]), int(clen[i])
                if cl < bl:
                    need(4); out[pos:pos + 4] = np.frombuffer(struct.pack("<I", cl), dtype=np.uint8); pos += 4
                    need(cl); out[pos:pos + cl] = scratch[do[i]:do[i] + cl]; pos += cl
                else:   # storing the block is no larger than its compressed form :120-125
                    need(4); out[pos:pos + 4] = np.frombuffer(struct.pack("<I", bl | UNCOMPRESSED_BLOCK_FLAG), dtype=np.uint8); pos += 4
                    need(bl); out[pos:pos + bl] = src[so[i]:so[i] + bl]; pos += bl
        need(4); out[pos:pos + 4] = 0; pos += 4                                     # end mark :130
        return pos

    def close(self):
        self._engine.close()


class _Block:
    __slots__ = ("pos", "length", "stored", "max_size", "checksum_pos", "frame")


class _Frame:
    __slots__ = ("first_block", "n_blocks", "content_checksum_pos", "has_content_size", "expected_content_size", "end_pos", "error")


class Lz4FrameCudaDecompressor:
    """Lz4FrameDecompressor (lz4/Lz4FrameDecompressor.java:21-31): all frames of the input, every block of a call in one batch."""

    BLOCKS_PER_BATCH_BYTES = 512 << 20      # scratch budget for one decode batch (each block gets a slot of its maximum size)

    def __init__(self, device=0):
        self._engine = BatchEngine(device)
        self._x = _Xxh32(self._engine._ctx)

    # ---- pass 1: the header walk (Lz4FrameCompression.java:146-346 without the block payloads)
    def _walk(self, inp, n_in):
        def u32(p):
            return int(inp[p]) | int(inp[p + 1]) << 8 | int(inp[p + 2]) << 16 | int(inp[p + 3]) << 24

        frames, blocks = [], []
        pos = 0
        error = None

        def fail(offset, text):
            return MalformedInputException(offset, text)

        while pos < n_in and error is None:
            if pos + 4 > n_in:
                error = fail(pos, "Truncated LZ4 frame: incomplete magic number"); break
            magic = u32(pos)
            if magic == MAGIC:
                fr = _Frame()
                fr.first_block, fr.n_blocks, fr.error = len(blocks), 0, None
                fr.content_checksum_pos, fr.expected_content_size, fr.end_pos, fr.has_content_size = -1, -1, -1, False
                frames.append(fr)
                d0 = pos + 4
                if d0 + 2 > n_in:
                    error = fail(d0, "Truncated LZ4 frame header"); break
                flg, bd = int(inp[d0]), int(inp[d0 + 1])
                version = (flg >> 6) & 3
                if version != 1:
                    error = fail(d0, f"Unsupported LZ4 frame version: {version}"); break
                if (flg & FLG_RESERVED_MASK) or (bd & BD_RESERVED_MASK):
                    error = fail(d0, "Corrupt LZ4 frame: reserved bits in the frame descriptor must be zero"); break
                if not flg & FLG_BLOCK_INDEPENDENCE:
                    error = fail(d0, "LZ4 frames with linked blocks are not supported"); break
                if flg & FLG_DICTIONARY_ID:
                    error = fail(d0, "LZ4 frames with a dictionary are not supported"); break
                bmax = block_maximum_size((bd >> 4) & 7)
                if bmax < 0:
                    error = fail(d0 + 1, "Invalid LZ4 frame block maximum size"); break
                p = d0 + 2
                has_size = bool(flg & FLG_CONTENT_SIZE)
                if p + (8 if has_size else 0) + 1 > n_in:
                    error = fail(p, "Truncated LZ4 frame header"); break
                if has_size:
                    fr.has_content_size = True
                    fr.expected_content_size = int.from_bytes(inp[p:p + 8].tobytes(), "little", signed=True)
                    p += 8
                if int(inp[p]) != (self._x.hash(inp, d0, p - d0) >> 8) & 0xFF:
                    error = fail(p, "Corrupt LZ4 frame: invalid header checksum"); break
                p += 1
                block_checksum = bool(flg & FLG_BLOCK_CHECKSUM)
                while True:
                    if p + 4 > n_in:
                        error = fail(p, "Truncated LZ4 frame: missing block size"); break
                    hdr = u32(p)
                    p += 4
                    if hdr == 0:
                        break
                    b = _Block()
                    b.stored, b.length, b.pos, b.max_size, b.frame = bool(hdr & UNCOMPRESSED_BLOCK_FLAG), hdr & BLOCK_SIZE_MASK, p, bmax, fr
                    if b.length > bmax or p + b.length > n_in:
                        error = fail(p, "Truncated LZ4 frame: block extends past end of input"); break
                    b.checksum_pos = -1
                    blocks.append(b)
                    fr.n_blocks += 1
                    if block_checksum:
                        b.checksum_pos = p + b.length
                        if b.checksum_pos + 4 > n_in:
                            b.checksum_pos = -2          # the block is decoded first, then "missing block checksum" (:280-283)
                            error = fail(p + b.length, "Truncated LZ4 frame: missing block checksum"); break
                    p += b.length + (4 if block_checksum else 0)
                if error is not None:
                    break
                if flg & FLG_CONTENT_CHECKSUM:
                    if p + 4 > n_in:
                        error = fail(p, "Truncated LZ4 frame: missing content checksum"); break
                    fr.content_checksum_pos = p
                    p += 4
                fr.end_pos = p
                pos = p
            elif (magic & SKIPPABLE_MAGIC_MASK) == SKIPPABLE_MAGIC:      # skipFrame :320-335
                sp = pos + 4
                if sp + 4 > n_in:
                    error = fail(sp, "Truncated LZ4 skippable frame: missing frame size"); break
                end = sp + 4 + u32(sp)
                if end > n_in:
                    error = fail(sp, "Truncated LZ4 skippable frame"); break
                pos = end
            else:
                error = fail(pos, "Invalid LZ4 frame magic number"); break
        if error is not None and frames and frames[-1].end_pos < 0:
            frames[-1].error = error          # found inside the last frame: reported after the blocks it had collected until then
            error = None
        return frames, blocks, error

    def decompress(self, input, inputOffset, inputLength, output, outputOffset, maxOutputLength):
        """Lz4FrameCompression.decompress (Lz4FrameCompression.java:135-180)."""
        src = _as_array(input)
        dst = _as_array(output, writable=True)
        _verify_range(src, inputOffset, inputLength)
        _verify_range(dst, outputOffset, maxOutputLength)
        inp = src[inputOffset:inputOffset + inputLength]
        out = dst[outputOffset:outputOffset + maxOutputLength]
        if inputLength < HEADER_SIZE:
            raise MalformedInputException(0, "Input is too short to be an LZ4 frame")
        frames, blocks, walk_error = self._walk(inp, inputLength)

        # ---- pass 2: every compressed block of the call, in batches (a block gets a scratch slot of its frame's maximum
        # size + 16: what a block decodes to is only known afterwards, and more than the maximum must be detected :266-268)
        comp = [i for i, b in enumerate(blocks) if not b.stored]
        dec_len = {}
        dec_status = {}
        dec_data = {}
        i0 = 0
        while i0 < len(comp):
            slot_bytes, i1 = 0, i0
            while i1 < len(comp) and (i1 == i0 or slot_bytes + blocks[comp[i1]].max_size + 16 <= self.BLOCKS_PER_BATCH_BYTES):
                slot_bytes += blocks[comp[i1]].max_size + 16
                i1 += 1
            ids = comp[i0:i1]
            so = np.array([blocks[i].pos for i in ids], dtype=np.int64)
            sl = np.array([blocks[i].length for i in ids], dtype=np.int64)
            dc = np.array([blocks[i].max_size + 16 for i in ids], dtype=np.int64)
            do = np.concatenate([[0], np.cumsum(dc)[:-1]]).astype(np.int64)
            scratch = np.empty(int(dc.sum()), dtype=np.uint8)
            olen, st = self._engine.run_host(N.OP_LZ4_DECOMPRESS, inp, so, sl, scratch, do, dc)
            for k, i in enumerate(ids):
                dec_len[i], dec_status[i] = int(olen[k]), int(st[k])
                dec_data[i] = scratch[do[k]:do[k] + max(int(olen[k]), 0)] if st[k] == 0 else None
            i0 = i1

        # block checksums: one XXH32 batch over the stored bytes of every block that carries one
        with_sum = [i for i, b in enumerate(blocks) if b.checksum_pos >= 0]
        sums = dict(zip(with_sum, self._x.hash_many(inp, [blocks[i].pos for i in with_sum], [blocks[i].length for i in with_sum])))

        # ---- pass 3: in the order of the Java loop -- place the blocks, report the first error it would have met
        out_pos = 0
        for fr in frames:
            frame_start = out_pos
            for i in range(fr.first_block, fr.first_block + fr.n_blocks):
                b = blocks[i]
                if b.stored:
                    if out_pos + b.length > maxOutputLength:
                        raise MalformedInputException(out_pos, "Output buffer too small")
                    out[out_pos:out_pos + b.length] = inp[b.pos:b.pos + b.length]
                    out_pos += b.length
                else:
                    if dec_status[i] != 0:
                        code, reason = dec_status[i] & 0xFF, dec_status[i] >> 8
                        text = N.lib().acc_reason_text(reason).decode()
                        if code == N.E_DST_TOO_SMALL:                   # more than the slot: more than the block maximum
                            raise MalformedInputException(b.pos, "Corrupt LZ4 frame: decompressed block exceeds maximum block size")
                        raise MalformedInputException(dec_len[i], text)   # the block decoder's own report (offset inside the block)
                    n = dec_len[i]
                    if out_pos + n > maxOutputLength:
                        raise MalformedInputException(out_pos, "Output buffer too small")
                    if n > b.max_size:
                        raise MalformedInputException(b.pos, "Corrupt LZ4 frame: decompressed block exceeds maximum block size")
                    out[out_pos:out_pos + n] = dec_data[i]
                    out_pos += n
                if b.checksum_pos == -2:
                    raise fr.error
                if b.checksum_pos >= 0:
                    expected = int.from_bytes(inp[b.checksum_pos:b.checksum_pos + 4].tobytes(), "little")
                    if expected != int(sums[i]):
                        raise MalformedInputException(b.checksum_pos, "Corrupt LZ4 frame: invalid block checksum")
            if fr.error is not None:
                raise fr.error
            content_length = out_pos - frame_start
            if fr.content_checksum_pos >= 0:
                expected = int.from_bytes(inp[fr.content_checksum_pos:fr.content_checksum_pos + 4].tobytes(), "little")
                if expected != self._x.hash(out, frame_start, content_length):
                    raise MalformedInputException(fr.content_checksum_pos, "Corrupt LZ4 frame: invalid content checksum")
            if fr.has_content_size and content_length != fr.expected_content_size:
                raise MalformedInputException(fr.end_pos, "Corrupt LZ4 frame: content size does not match frame header")
        if walk_error is not None:
            raise walk_error
        return out_pos

    def close(self):
        self._engine.close()
