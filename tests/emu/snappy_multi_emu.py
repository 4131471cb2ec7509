"""Lane-level emulation (numpy, 32 lanes) of the multi-element Snappy decode step in snappy.cu, checked against the
oracle on corpus blocks and corrupted streams.  Development aid (see lz4_multiseq_emu.py).  python tests/emu/snappy_multi_emu.py (also run by tests/test_step_emulators.py)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

LANE = np.arange(32, dtype=np.int64)
STAT = {"multi": 0, "elems": 0, "pair": 0, "slow": 0, "rounds": 0}
MULTI = True


def shfl(v, idx):
    return v[np.asarray(idx) & 31]


def op_entry(op):
    kind, hi = op & 3, op >> 2
    if kind == 0:
        return hi + 1 if hi < 60 else (((hi - 59) << 11) | 1)
    if kind == 1:
        return (1 << 11) | ((op >> 5) << 8) | (4 + (hi & 7))
    if kind == 2:
        return (2 << 11) | (hi + 1)
    return (4 << 11) | (hi + 1)


def decode(inp, out_cap):
    """returns (reason or 0, out_len_or_offset, out)"""
    n0 = len(inp)
    src0 = np.frombuffer(inp, dtype=np.uint8).astype(np.int64)
    out = np.zeros(out_cap + 128, dtype=np.int64)
    # preamble
    result = 0; nread = 0; shift = 0
    while True:
        if nread >= n0:
            return ("TRUNCATED", n0 - nread, out)
        b = int(src0[nread]); nread += 1
        result |= (b & 0x7f) << shift
        if not (b & 0x80):
            break
        if shift == 28:
            return ("VARINT_HIGHBIT", nread, out)
        shift += 7
    result &= 0xffffffff
    if result >= 0x80000000:
        return ("NEG_LENGTH", 0, out)
    expected = result
    if expected > out_cap:
        return ("LEN_GT_CAP", 0, out)
    src = src0[nread:]
    n = n0 - nread
    fast_output_limit = out_cap - 8
    ip = op = 0
    srcp = np.concatenate([src, np.zeros(64, dtype=np.int64)])
    while ip < n:
        if ip + 32 <= n:
            vb = srcp[ip + LANE]
            if MULTI and op + 32 <= out_cap:
                kind = vb & 3; hi = vb >> 2
                b1 = shfl(vb, LANE + 1); b2 = shfl(vb, LANE + 2)
                outn = np.where(kind == 0, hi + 1, np.where(kind == 1, 4 + (hi & 7), hi + 1))
                adv = np.where(kind == 0, hi + 2, np.where(kind == 1, 2, 3))
                off = np.where(kind == 1, ((vb >> 5) << 8) | b1, np.where(kind == 2, b1 | (b2 << 8), 0))
                usable = (kind != 3) & ~((kind == 0) & (hi >= 60)) & ~((kind != 0) & (off == 0)) & (LANE + adv <= 32) & (outn <= 32)
                n_l = np.where(usable, outn, 127)
                A = n_l | (adv << 8) | ((kind == 0).astype(np.int64) << 14)
                a0 = int(A[0]); o0 = int(off[0])
                nn0 = a0 & 127
                lit0 = (a0 >> 14) & 1
                if nn0 <= 32 and o0 <= op:
                    pos = [0, 0, 0, 0]; base = [0, 0, 0, 0]; lits = [lit0, 0, 0, 0]; offs = [o0, 1, 1, 1]
                    e = nn0; nx = (a0 >> 8) & 63; cnt = 1
                    while cnt < 4 and nx < 32:
                        ak = int(A[nx]); ok = int(off[nx]); nk = ak & 127; lk = (ak >> 14) & 1
                        if not (e + nk <= 32 and ok <= op + e):
                            break
                        pos[cnt] = nx; base[cnt] = e; lits[cnt] = lk; offs[cnt] = ok
                        e += nk; nx += (ak >> 8) & 63; cnt += 1
                    if cnt >= 2:
                        k = np.zeros(32, dtype=np.int64)
                        for q in range(1, cnt):
                            k += (LANE >= base[q]).astype(np.int64)
                        sk = np.array(pos)[k]; bk = np.array(base)[k]; lk = np.array(lits)[k]; fk = np.array(offs)[k]
                        t = LANE - bk
                        is_lit = lk == 1
                        lit = shfl(vb, sk + 1 + t)
                        m = np.where(~is_lit & (t >= fk), t % np.maximum(fk, 1), t)
                        srel = bk - fk + m
                        active = LANE < e
                        val = np.where(is_lit, lit, 0)
                        pending = active & ~is_lit
                        frommem = pending & (srel < 0)
                        for j in np.flatnonzero(frommem):
                            assert op + srel[j] >= 0
                            val[j] = out[op + srel[j]]
                        pending = pending & ~frommem
                        rounds = 0
                        while pending.any():
                            w = shfl(val | (pending.astype(np.int64) << 8), srel)
                            got = pending & ((w & 0x100) == 0)
                            val = np.where(got, w & 0xff, val)
                            pending = pending & ~got
                            rounds += 1
                            assert rounds <= 5
                        for j in np.flatnonzero(active):
                            out[op + j] = val[j]
                        STAT["multi"] += 1; STAT["elems"] += cnt; STAT["rounds"] += rounds
                        ip += nx; op += e
                        continue
            # ---- pair path (existing kernel fast path)
            t0 = int(vb[0]); L = 0; p = 0; ok = True
            if (t0 & 3) == 0:
                nn = t0 >> 2
                if nn <= 26:
                    L = nn + 1; p = 1 + L
                else:
                    ok = False
            if ok:
                tag = int(vb[p]); c1 = int(vb[(p + 1) & 31]); c2 = int(vb[(p + 2) & 31])
                kind = tag & 3; clen = 0; coff = 1; adv = p
                if kind == 1:
                    clen = 4 + ((tag >> 2) & 7); coff = ((tag >> 5) << 8) | c1; adv = p + 2
                elif kind == 2:
                    clen = (tag >> 2) + 1; coff = c1 | (c2 << 8); adv = p + 3
                elif L == 0:
                    ok = False
                total = L + clen
                if ok and coff != 0 and coff <= op + L and op + total <= out_cap:
                    for j in range(L):
                        out[op + j] = vb[1 + j]
                    for j in range(clen):
                        out[op + L + j] = out[op + L + j - coff]
                    STAT["pair"] += 1
                    ip += adv; op += total
                    continue
        STAT["slow"] += 1
        opc = int(srcp[ip]); ip += 1
        entry = op_entry(opc)
        tb = entry >> 11
        if not (ip + 4 < n):
            if ip + tb > n:
                return ("NONE", ip, out)
        trailer = 0
        for i in range(tb - 1, -1, -1):
            trailer = (trailer << 8) | int(srcp[ip + i])
        if trailer >= 0x80000000:
            return ("NONE", ip, out)
        ip += tb
        length = entry & 0xff
        if (opc & 3) == 0:
            ll = (length + trailer) & 0xffffffff
            if ll >= 0x80000000:
                return ("NONE", ip, out)
            lol = op + ll
            if lol > fast_output_limit or ip + ll > n - 8:
                if lol > out_cap or ip + ll > n:
                    return ("NONE", ip, out)
            out[op:op + ll] = srcp[ip:ip + ll]
            ip += ll; op = lol
        else:
            moff = ((entry & 0x700) + trailer) & 0xffffffff
            if moff >= 0x80000000 or moff == 0:
                return ("NONE", ip, out)
            if moff > op or op + length > out_cap:
                return ("NONE", ip, out)
            for i in range(length):
                out[op + i] = out[op + i - moff]
            op += length
    if expected != op:
        return ("LEN_MISMATCH", 0, out)
    return (0, expected, out)


REASON = {0: "NONE", 7: "TRUNCATED", 8: "VARINT_HIGHBIT", 9: "NEG_LENGTH", 10: "LEN_GT_CAP", 11: "LEN_MISMATCH"}


def main(o=None, n_cases=40):
    import benchdata
    if o is None:
        from oracle.pyoracle import Oracle
        o = Oracle()
    blob = np.fromfile(os.path.join(benchdata.ROOT, "tests", "golden", "silesia_sample.bin"), dtype=np.uint8)
    rng = np.random.default_rng(7)
    starts = rng.integers(0, len(blob) - 16384, size=n_cases)
    cases = []
    for st in starts:
        sz = int(rng.choice([300, 2000, 8192, 16384]))
        cases.append(bytes(blob[st:st + sz]))
    cases += [b"a" * 5000, bytes(range(256)) * 20, b"ab" * 3000, b"abc" * 2000 + b"xyz" * 700, (b"0123456789abcdefg" * 400)]
    checked = bad = 0
    for raw in cases:
        c = o.compress("snappy", raw)
        variants = [(c, len(raw)), (c, len(raw) + 100), (c, len(raw) - 1), (c, len(raw) + 11)]
        for _ in range(6):
            cc = bytearray(c)
            for _ in range(int(rng.integers(1, 3))):
                cc[int(rng.integers(0, len(cc)))] ^= int(rng.integers(1, 256))
            variants.append((bytes(cc), len(raw) + int(rng.integers(0, 64))))
        variants.append((c[:len(c) // 2], len(raw)))
        for comp, cap in variants:
            st, ln, out = decode(comp, cap)
            r, off, eout = o.decompress_raw("snappy", comp, cap)
            if r >= 0:
                exp = eout[:r].tobytes(); eo = (0, r)
            else:
                exp = None; eo = (REASON.get((-r) >> 8, str((-r) >> 8)), off)
            checked += 1
            if (st, ln) != eo or (exp is not None and bytes(out[:ln].astype(np.uint8)) != exp):
                bad += 1
                print("MISMATCH", (st, ln), eo, len(comp), cap)
    print("checked", checked, "bad", bad, STAT, "elems/multi", STAT["elems"] / max(1, STAT["multi"]))
    return checked, bad, dict(STAT)


if __name__ == "__main__":
    main()
