"""Hadoop block streams over the batched GPU block codecs (the LZ4 and Snappy block streams).

The reference's streams (lz4/Lz4HadoopOutputStream.java:36-131, lz4/Lz4HadoopInputStream.java:38-163 and their Snappy twins)
speak Hadoop's block format -- per block `[BE32 uncompressed length]`, then chunks `[BE32 compressed length][bytes]` until
the block's length is covered -- and hand ONE chunk at a time to the block codec (256 KiB buffers,
lz4/Lz4HadoopStreams.java:29).  These adapters keep the format and the chunk geometry (a chunk is cut exactly where the
reference cuts it: buffer size minus the codec's overhead allowance, Lz4HadoopOutputStream.java:128-131 /
SnappyHadoopOutputStream.java:128-131) but collect many chunks and submit them as one batch:

  output   chunks are buffered until `batch_chunks` are complete (or flush / finish / close), compressed by one acc_batch
           call and written as `[BE32 raw][BE32 comp][bytes]` each, one chunk per block like the reference writes;
  input    the reader walks the headers ahead, collects up to `batch_chunks` chunks, decodes them as one batch and serves
           reads from the result.  What a chunk decodes to is not in an LZ4 chunk, so for LZ4 the walk assumes what the
           reference writer produces (one chunk per block) and re-reads chunk by chunk behind the first block that holds
           several; a Snappy chunk announces its length itself.  Whatever the walk runs into (end of stream inside a
           chunk, a truncated length) is raised only after the chunks in front of it have been delivered, as a sequential
           reader would.

Python file objects play the role of java.io streams: `out.write(bytes)`, `in.read(n)`.  No CPU fallback.
"""
import ctypes as C
import io

import numpy as np

from . import _native as N
from .api import BatchEngine, MalformedInputException

DEFAULT_BUFFER_SIZE = 256 * 1024                 # Lz4HadoopStreams.java:29, SnappyHadoopStreams.java:29


def lz4_overhead(size):
    return max(int(size * 0.01), 10)             # Lz4HadoopOutputStream.java:128-131


def snappy_overhead(size):
    return size // 6 + 32                        # SnappyHadoopOutputStream.java:128-131


def _be32(v):
    return bytes(((v >> 24) & 0xFF, (v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF))


class _HadoopCudaOutputStream(io.RawIOBase):
    codec = None

    def __init__(self, out, buffer_size=DEFAULT_BUFFER_SIZE, batch_chunks=256, device=0, engine=None):
        super().__init__()
        self._engine = engine if engine is not None else BatchEngine(device)
        self._out = out
        self._chunk = buffer_size - self._overhead(buffer_size)        # inputMaxSize
        self._batch = max(1, batch_chunks)
        self._pending = bytearray()
        self._bound = int(getattr(N.lib(), f"acc_{self.codec}_compress_bound")(self._chunk))

    def writable(self):
        return True

    def write(self, data):
        self._pending += data
        if len(self._pending) >= self._chunk * self._batch:
            self._emit(len(self._pending) // self._chunk * self._chunk)
        return len(data)

    def _emit(self, nbytes):
        """compresses pending[:nbytes] -- whole chunks, or everything at finish() -- as one batch and writes the blocks"""
        if nbytes == 0:
            return
        src = np.frombuffer(bytes(self._pending[:nbytes]), dtype=np.uint8)
        del self._pending[:nbytes]
        n = (nbytes + self._chunk - 1) // self._chunk
        so = np.arange(n, dtype=np.int64) * self._chunk
        sl = np.minimum(self._chunk, nbytes - so)
        do = np.arange(n, dtype=np.int64) * self._bound
        dst = np.empty(n * self._bound, dtype=np.uint8)
        op = N.OP_LZ4_COMPRESS if self.codec == "lz4" else N.OP_SNAPPY_COMPRESS
        clen, st = self._engine.run_host(op, src, so, sl, dst, do, np.full(n, self._bound, dtype=np.int64))
        if (st != 0).any():
            raise RuntimeError(f"block compression failed: status {int(st[st != 0][0]):#x}")
        for i in range(n):                                             # writeNextChunk :107-117
            self._out.write(_be32(int(sl[i])) + _be32(int(clen[i])))
            self._out.write(dst[do[i]:do[i] + int(clen[i])].tobytes())

    def flush(self):
        # the reference has written every complete chunk by the time flush() is called; so has this stream afterwards
        if not self.closed:
            self._emit(len(self._pending) // self._chunk * self._chunk)
            if hasattr(self._out, "flush"):
                self._out.flush()

    def finish(self):
        """HadoopOutputStream.finish (:81-88): the partial chunk goes out too"""
        self._emit(len(self._pending))

    def close(self):
        if not self.closed:
            try:
                self.finish()
            finally:
                super().close()
                if hasattr(self._out, "close"):
                    self._out.close()


class Lz4HadoopCudaOutputStream(_HadoopCudaOutputStream):
    codec = "lz4"
    _overhead = staticmethod(lz4_overhead)


class SnappyHadoopCudaOutputStream(_HadoopCudaOutputStream):
    codec = "snappy"
    _overhead = staticmethod(snappy_overhead)


class _Replay:
    """The input with a memory: everything read since mark() can be read again after rewind() -- the LZ4 reader walks ahead on
    an assumption (one chunk per block) that only decoding can confirm."""

    def __init__(self, inner):
        self._inner, self._buf, self._pos = inner, bytearray(), 0

    def mark(self):
        del self._buf[:self._pos]
        self._pos = 0

    def tell(self):
        return self._pos

    def rewind(self, pos):
        self._pos = pos

    def read(self, n):
        while len(self._buf) - self._pos < n:
            more = self._inner.read(n - (len(self._buf) - self._pos))
            if not more:
                break
            self._buf += more
        d = bytes(self._buf[self._pos:self._pos + n])
        self._pos += len(d)
        return d

    def close(self):
        if hasattr(self._inner, "close"):
            self._inner.close()


class _HadoopCudaInputStream(io.RawIOBase):
    codec = None

    def __init__(self, inp, buffer_size=DEFAULT_BUFFER_SIZE, batch_chunks=64, device=0, engine=None):
        super().__init__()
        self._engine = engine if engine is not None else BatchEngine(device)
        self._in = _Replay(inp)
        self._batch = max(1, batch_chunks)
        self._ready = bytearray()        # decoded, not yet delivered
        self._pending_error = None       # raised once `_ready` is drained
        self._eof = False
        self._block_left = 0             # bytes of the current block not yet covered by decoded chunks
        self._sequential = False         # LZ4: a block of several chunks was seen -> chunk by chunk from here on

    def readable(self):
        return True

    # ---- the header walk (bufferCompressedData :100-128, readBigEndianInt :147-162, readInput :130-145)
    def _read_int(self):
        b = self._in.read(4)
        if len(b) == 0:
            return -1
        if len(b) < 4:
            raise IOError("Stream is truncated")
        return (b[0] << 24) + (b[1] << 16) + (b[2] << 8) + b[3]

    def _next_chunk(self):
        """-> (compressed bytes, block bytes left before this chunk) or None at the end of the stream"""
        while self._block_left == 0:
            v = self._read_int()
            if v == -1:
                return None
            self._block_left = v
        n = self._read_int()
        if n == -1:
            return None
        data = self._in.read(n)
        if len(data) < n:
            raise EOFError("encountered EOF while reading block data")
        return data, self._block_left

    def _declared_length(self, chunk, block_left):
        if self.codec == "snappy":       # the chunk's own preamble (SnappyRawDecompressor.readUncompressedLength)
            off = C.c_int64(0)
            src = np.frombuffer(chunk, dtype=np.uint8) if chunk else np.zeros(1, dtype=np.uint8)
            r = N.lib().acc_snappy_uncompressed_length(src.ctypes.data, len(chunk), C.byref(off))
            return int(r) if r >= 0 else block_left
        return block_left                # LZ4: assume the reference writer's one chunk per block; _fill() checks

    def _fill(self):
        chunks, caps, after = [], [], []
        limit = 1 if self._sequential else self._batch
        self._in.mark()
        try:
            while len(chunks) < limit:
                nxt = self._next_chunk()
                if nxt is None:
                    self._eof = True
                    break
                chunk, left = nxt
                chunks.append(chunk)
                caps.append(left)
                after.append(self._in.tell())
                self._block_left = max(left - self._declared_length(chunk, left), 0)   # provisional for LZ4, exact for Snappy
        except (IOError, EOFError) as e:
            self._pending_error = e
            self._eof = True
        if not chunks:
            return
        lens = np.array([len(c) for c in chunks], dtype=np.int64)
        so = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        src = np.frombuffer(b"".join(chunks), dtype=np.uint8) if lens.sum() else np.zeros(1, dtype=np.uint8)
        dc = np.array(caps, dtype=np.int64)
        do = np.concatenate([[0], np.cumsum(dc)[:-1]]).astype(np.int64)
        dst = np.empty(max(int(dc.sum()), 1), dtype=np.uint8)
        op = N.OP_LZ4_DECOMPRESS if self.codec == "lz4" else N.OP_SNAPPY_DECOMPRESS
        olen, st = self._engine.run_host(op, src, so, lens, dst, do, dc)
        for i in range(len(chunks)):
            if st[i] != 0:
                self._pending_error = MalformedInputException(int(olen[i]), N.lib().acc_reason_text(int(st[i]) >> 8).decode())
                self._eof = True
                return
            got = int(olen[i])
            self._ready += dst[do[i]:do[i] + got].tobytes()
            if self.codec == "lz4" and got < caps[i]:
                # a block of several chunks: what was walked behind this chunk was framed on a wrong assumption (and whatever
                # the walk ran into there means nothing).  Go back to the byte behind this chunk and go on chunk by chunk.
                self._in.rewind(after[i])
                self._block_left = caps[i] - got
                self._sequential = True
                self._eof = False
                self._pending_error = None
                return

    def readinto(self, b):
        while not self._ready and not self._eof:
            self._fill()
        if not self._ready:
            if self._pending_error is not None:
                e, self._pending_error = self._pending_error, None
                raise e
            return 0
        n = min(len(b), len(self._ready))
        b[:n] = self._ready[:n]
        del self._ready[:n]
        return n

    def close(self):
        if not self.closed:
            super().close()
            self._in.close()


class Lz4HadoopCudaInputStream(_HadoopCudaInputStream):
    codec = "lz4"


class SnappyHadoopCudaInputStream(_HadoopCudaInputStream):
    codec = "snappy"
