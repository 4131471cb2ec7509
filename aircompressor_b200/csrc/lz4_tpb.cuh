// lz4_tpb.cuh -- LZ4 block decoder, one THREAD per block ("tpb"), written as plain scalar code over global memory.
//
// Rationale (profiles/README.md): the warp-per-block decoder spends ~110 warp-instructions per sequence because
// 30 of 32 lanes idle in every instruction.  Here every lane decodes its own block, so one warp-instruction advances
// 32 sequences.  All loads and stores are 16-byte aligned vector accesses: unaligned sources are assembled from two
// aligned chunks with byte funnel shifts, the output is appended through a 16-byte "pending chunk" register that is
// re-stored after every append (garbage beyond the write position is harmless inside the block's own output region,
// exactly like the wild copies of the CPU implementations; the first and last bytes of a block use exact byte stores
// so neighbouring blocks are never touched).
//
// The routine is an optimistic decoder: it handles the valid common shape of a block and returns kFallback for
// anything else (malformed input, capacity corner cases of Lz4RawDecompressor.java:82-96,168-171).  The caller then
// re-decodes that block with the exact decoder of lz4_decode_v1.cuh, so results, reject decisions and error offsets
// are those of the Java decoder.  The same source compiles for the host (tests/ compile it with g++ to check it
// against the oracle without a GPU).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TPB_HD __host__ __device__ __forceinline__
#else
#define TPB_HD static inline
#endif

namespace lz4tpb {

enum { kOk = 0, kFallback = 1 };

struct U128 { uint64_t lo, hi; };

TPB_HD U128 ld16a(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    U128 r;
    r.lo = (uint64_t) v.x | ((uint64_t) v.y << 32);
    r.hi = (uint64_t) v.z | ((uint64_t) v.w << 32);
    return r;
#else
    U128 r;
    memcpy(&r, p, 16);
    return r;
#endif
}

TPB_HD void st16a(uint8_t *p, U128 v)
{
#if defined(__CUDA_ARCH__)
    uint4 w;
    w.x = (uint32_t) v.lo; w.y = (uint32_t) (v.lo >> 32); w.z = (uint32_t) v.hi; w.w = (uint32_t) (v.hi >> 32);
    *reinterpret_cast<uint4 *>(p) = w;
#else
    memcpy(p, &v, 16);
#endif
}

// bytes [sh, sh + 16) of the 32-byte string a|b, sh in 0..15
TPB_HD U128 shr_bytes(U128 a, U128 b, unsigned sh)
{
    const unsigned s = sh * 8;
    U128 r;
    if (s == 0) return a;
    if (s < 64) { r.lo = (a.lo >> s) | (a.hi << (64 - s)); r.hi = (a.hi >> s) | (b.lo << (64 - s)); }
    else if (s == 64) { r.lo = a.hi; r.hi = b.lo; }
    else { const unsigned t = s - 64; r.lo = (a.hi >> t) | (b.lo << (64 - t)); r.hi = (b.lo >> t) | (b.hi << (64 - t)); }
    return r;
}

// a shifted left by sh bytes (sh in 0..15), zero filled
TPB_HD U128 shl_bytes(U128 a, unsigned sh)
{
    const unsigned s = sh * 8;
    U128 r;
    if (s == 0) return a;
    if (s < 64) { r.lo = a.lo << s; r.hi = (a.hi << s) | (a.lo >> (64 - s)); }
    else if (s == 64) { r.lo = 0; r.hi = a.lo; }
    else { r.lo = 0; r.hi = a.lo << (s - 64); }
    return r;
}

// keep the low n bytes (n in 0..16)
TPB_HD U128 low_bytes(U128 a, unsigned n)
{
    U128 r;
    if (n >= 16) return a;
    if (n == 0) { r.lo = 0; r.hi = 0; }
    else if (n < 8) { r.lo = a.lo & ((1ull << (8 * n)) - 1); r.hi = 0; }
    else if (n == 8) { r.lo = a.lo; r.hi = 0; }
    else { r.lo = a.lo; r.hi = a.hi & ((1ull << (8 * (n - 8))) - 1); }
    return r;
}

// 16 bytes at an arbitrary address (the two aligned chunks that contain them are read)
TPB_HD U128 load16u(const uint8_t *p)
{
    const unsigned sh = (unsigned) ((uintptr_t) p & 15);
    const uint8_t *q = p - sh;
    const U128 x = ld16a(q);
    if (sh == 0) return x;
    return shr_bytes(x, ld16a(q + 16), sh);
}

// output cursor: `chunk` = 16-byte aligned address of the chunk that contains the write position, `pend` = that
// chunk's bytes (those below the write position are valid)
struct OutCursor {
    uint8_t *chunk;
    U128 pend;
};

// append n (1..16) bytes of `data` at absolute address `cur`; may write up to `cur - (cur & 15) + 32`
TPB_HD void append16(OutCursor &o, uint8_t *cur, U128 data, unsigned n)
{
    const unsigned sh = (unsigned) ((uintptr_t) cur & 15);
    U128 a = low_bytes(o.pend, sh);
    const U128 d = shl_bytes(data, sh);
    a.lo |= d.lo; a.hi |= d.hi;
    st16a(o.chunk, a);
    if (sh + n >= 16) {
        U128 zero; zero.lo = 0; zero.hi = 0;
        const U128 spill = sh ? shr_bytes(data, zero, 16 - sh) : zero;
        if (sh + n > 16) st16a(o.chunk + 16, spill);
        o.chunk += 16;
        o.pend = spill;
    }
    else {
        o.pend = a;
    }
}

// copies exactly n bytes dst[0..n) = src[0..n); regions do not overlap, or dst - src >= 16 (forward LZ77 copy)
TPB_HD void copy_exact(uint8_t *dst, const uint8_t *src, uint32_t n)
{
    if (n >= 48) {
        const uint32_t head = (uint32_t) ((16 - ((uintptr_t) dst & 15)) & 15);
        for (uint32_t k = 0; k < head; k++) dst[k] = src[k];
        dst += head; src += head; n -= head;
        while (n >= 16) { st16a(dst, load16u(src)); dst += 16; src += 16; n -= 16; }
    }
    for (uint32_t k = 0; k < n; k++) dst[k] = src[k];
}

// exact byte-wise decoding of whole sequences starting at (ip, op) until op >= op_stop (checked between sequences) or
// the final literal run is done.  Returns 0 = stopped at op_stop, 1 = block finished, 2 = fallback.
TPB_HD int decode_bytes(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, uint32_t &ip, uint32_t &op, uint32_t op_stop)
{
    while (ip < in_len) {
        if (op >= op_stop) return 0;
        const uint32_t token = in[ip];
        uint32_t p = ip + 1, ll = token >> 4;
        if (ll == 15) {
            if (p >= in_len) return 2;
            uint32_t v;
            do { v = in[p++]; ll += v; } while (v == 255 && p + 15 < in_len);
        }
        const uint64_t lit_end = (uint64_t) p + ll, lit_out = (uint64_t) op + ll;
        if (lit_out + 12 > out_cap || lit_end + 8 > in_len) {
            // final literal run (Lz4RawDecompressor.java:82-96)
            if (lit_out > out_cap || lit_end != in_len) return 2;
            copy_exact(out + op, in + p, ll);
            op += ll;
            ip = in_len;
            return 1;
        }
        copy_exact(out + op, in + p, ll);
        op += ll;
        p += ll;
        const uint32_t off = (uint32_t) in[p] | ((uint32_t) in[p + 1] << 8);
        p += 2;
        if (off == 0 || off > op) return 2;
        uint32_t ml = token & 15;
        if (ml == 15) {
            uint32_t v;
            do { if (p + 5 > in_len) return 2; v = in[p++]; ml += v; } while (v == 255);
        }
        ml += 4;
        const uint64_t match_out = (uint64_t) op + ml;
        if (match_out + 12 > out_cap && match_out + 5 > out_cap) return 2;
        if (off >= 16) copy_exact(out + op, out + op - off, ml);
        else for (uint32_t k = 0; k < ml; k++) out[op + k] = out[op - off + k];
        op += ml;
        ip = p;
    }
    return 2;   // a valid block ends with a final literal run, never here
}

// Decodes one block.  kOk: *out_len set, bytes [0, *out_len) of out written, nothing outside [out, out + out_cap) touched.
TPB_HD int decode_block(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, uint32_t *out_len)
{
    if (in_len == 0 || out_cap == 0) return kFallback;
    uint32_t ip = 0, op = 0;
    // head: exact bytes until the write position's 16-byte chunk lies inside this block's output
    int r = decode_bytes(in, in_len, out, out_cap, ip, op, 32);
    if (r == 2) return kFallback;
    if (r == 1) { *out_len = op; return kOk; }

    OutCursor oc;
    {
        uint8_t *cur = out + op;
        oc.chunk = cur - ((uintptr_t) cur & 15);
        oc.pend = ld16a(oc.chunk);
    }
    U128 zero; zero.lo = 0; zero.hi = 0;

    // State machine: every iteration performs at most one parse and exactly one <= 16-byte copy step, so that on the GPU
    // the 32 lanes of a warp (32 different blocks) never wait for a lane that is inside a long literal run or match.
    //   rem      bytes left in the current copy (0 = parse the next sequence)
    //   in_match current copy is the match (else: literals, after which the match (m_len, m_off) follows)
    uint32_t rem = 0, m_len = 0, m_off = 0, lit_src = 0;
    bool in_match = false;
    U128 pat = zero;
    uint32_t pat_step = 16;
    for (;;) {
        if (rem == 0) {
            // ---- parse one sequence; ip / op are only advanced once it is known to fit the wide path ----
            if ((uint64_t) ip + 20 > in_len) break;
            const U128 w = load16u(in + ip);
            const uint32_t token = (uint32_t) (w.lo & 0xFF);
            uint32_t ll = token >> 4, ml = token & 15, lit_pos, off, next_ip;
            uint64_t lit_end;
            bool lit_in_w;
            if (ll <= 12 && ml != 15) {
                // token, literals and offset are all inside the 16 loaded bytes
                lit_pos = ip + 1;
                lit_end = (uint64_t) lit_pos + ll;
                off = (uint32_t) (shr_bytes(w, zero, 1 + ll).lo & 0xFFFF);
                ml += 4;
                next_ip = (uint32_t) lit_end + 2;
                lit_in_w = true;
            }
            else {
                uint32_t p = ip + 1;
                if (ll == 15) {
                    uint32_t v;
                    do { v = in[p++]; ll += v; } while (v == 255 && p + 15 < in_len);
                }
                lit_pos = p;
                lit_end = (uint64_t) p + ll;
                if (lit_end + 16 > in_len) break;                   // final run, or too close to the end: the byte tail decides
                const uint32_t le = (uint32_t) lit_end;
                off = (uint32_t) in[le] | ((uint32_t) in[le + 1] << 8);
                uint32_t q = le + 2;
                if (ml == 15) {
                    uint32_t v;
                    do { if (q + 5 > in_len) return kFallback; v = in[q++]; ml += v; } while (v == 255);
                }
                ml += 4;
                next_ip = q;
                lit_in_w = false;
            }
            // bounds of the wide path; they imply the Java decoder's non-final rules (Lz4RawDecompressor.java:82, :168-171)
            if (lit_end + 16 > in_len) break;
            if ((uint64_t) op + ll + ml + 64 > out_cap) break;
            if (off == 0 || off > op + ll) return kFallback;         // "offset outside destination buffer": exact decoder reports it
            ip = next_ip;
            m_len = ml;
            m_off = off;
            if (lit_in_w) {
                if (ll) { append16(oc, out + op, shr_bytes(w, zero, 1), ll); op += ll; }
                in_match = true;
                rem = ml;
            }
            else if (ll == 0) { in_match = true; rem = ml; }
            else { in_match = false; rem = ll; lit_src = lit_pos; }
            if (in_match && m_off < 16) {
                // overlapping match: replicate the m_off-byte pattern once, then append multiples of it
                pat = low_bytes(load16u(out + op - m_off), m_off);
                for (uint32_t have = m_off; have < 16; have *= 2) {
                    const U128 sp = shl_bytes(pat, have);
                    pat.lo |= sp.lo; pat.hi |= sp.hi;
                }
                pat_step = (16 / m_off) * m_off;
            }
            if (!lit_in_w && ll != 0) continue;   // long literal run: start copying in the next iterations
        }
        // ---- one copy step ----
        if (!in_match) {
            const uint32_t n = rem < 16 ? rem : 16;
            append16(oc, out + op, load16u(in + lit_src), n);
            op += n; lit_src += n; rem -= n;
            if (rem == 0) {
                in_match = true;
                rem = m_len;
                if (m_off < 16) {
                    pat = low_bytes(load16u(out + op - m_off), m_off);
                    for (uint32_t have = m_off; have < 16; have *= 2) {
                        const U128 sp = shl_bytes(pat, have);
                        pat.lo |= sp.lo; pat.hi |= sp.hi;
                    }
                    pat_step = (16 / m_off) * m_off;
                }
            }
        }
        else if (m_off >= 16) {
            const uint32_t n = rem < 16 ? rem : 16;
            append16(oc, out + op, load16u(out + op - m_off), n);
            op += n; rem -= n;
        }
        else {
            const uint32_t n = rem < pat_step ? rem : pat_step;
            append16(oc, out + op, pat, n);
            op += n; rem -= n;
        }
    }
    // tail: exact bytes (everything below op is already in memory)
    r = decode_bytes(in, in_len, out, out_cap, ip, op, 0xFFFFFFFFu);
    if (r != 1) return kFallback;
    *out_len = op;
    return kOk;
}

}  // namespace lz4tpb
