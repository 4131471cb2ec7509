"""Lane-level emulation (numpy, 32 lanes) of the multi-sequence LZ4 fast path in lz4_decode_v1.cuh, checked against the
oracle on corpus blocks and corrupted streams.  Development aid: validates the index arithmetic of the kernel on the CPU
before GPU time is spent.  python tests/emu/lz4_multiseq_emu.py (also run by tests/test_step_emulators.py)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

LANE = np.arange(32, dtype=np.int64)
STAT = {"iters": 0, "seqs": 0, "rounds": 0, "slow": 0}


def shfl(v, idx):
    return v[np.asarray(idx) & 31]


MEDIUM = True


def decode(inp, out_cap, KMAX=3):
    """returns (status_reason or 0, out_len_or_offset, out)"""
    n = len(inp)
    src = np.frombuffer(inp, dtype=np.uint8).astype(np.int64)
    out = np.zeros(out_cap + 64, dtype=np.int64)
    ip = op = 0
    if n == 0:
        return ("INPUT_EMPTY", 0, out)
    if out_cap == 0:
        if n == 1 and inp[0] == 0:
            return (0, 0, out)
        return ("ZERO_CAP", 0, out)
    fast_output_limit = out_cap - 8
    while ip < n:
        if ip + 40 <= n and op + 44 <= out_cap:
            vb = src[ip + LANE]
            ll_l = vb >> 4
            ml_l = vb & 15
            nxt_l = LANE + 3 + ll_l
            ok_l = (ll_l != 15) & (ml_l != 15) & (nxt_l <= 32)
            off_l = shfl(vb, LANE + 1 + ll_l) | (shfl(vb, LANE + 2 + ll_l) << 8)
            P = ll_l | ((ll_l + ml_l + 4) << 4) | (ok_l.astype(np.int64) << 10) | (off_l << 16)
            # uniform chain
            p0 = int(P[0])
            if (p0 >> 10) & 1:
                ll0, n0, off0 = p0 & 15, (p0 >> 4) & 63, p0 >> 16
                if off0 == 0 or off0 > op + ll0:
                    return ("OFFSET_OUTSIDE", ip + ll0 + 3, out)
                s = [0, 0, 0]; base = [0, 0, 0]; lls = [ll0, 0, 0]; offs = [off0, 1, 1]
                cnt = 1
                e = n0           # output bytes so far
                nx = 3 + ll0     # next token position in the window
                while cnt < KMAX:
                    if nx >= 32:
                        break
                    pk = int(P[nx & 31])
                    llk, nk, offk = pk & 15, (pk >> 4) & 63, pk >> 16
                    if not ((pk >> 10) & 1) or e + nk > 32 or offk == 0 or offk > op + e + llk:
                        break
                    s[cnt] = nx; base[cnt] = e; lls[cnt] = llk; offs[cnt] = offk
                    e += nk; nx += 3 + llk; cnt += 1
                # per lane
                k = np.zeros(32, dtype=np.int64)
                for q in range(1, cnt):
                    k += (LANE >= base[q]).astype(np.int64)
                sk = np.array(s)[k]; bk = np.array(base)[k]; lk = np.array(lls)[k]; ok_ = np.array(offs)[k]
                t = LANE - bk
                is_lit = t < lk
                lit = shfl(vb, sk + 1 + t)
                m = t - lk
                m = np.where(m >= ok_, m % ok_, m)
                srel = bk + lk - ok_ + m
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
                    assert rounds <= 4
                for j in np.flatnonzero(active):
                    out[op + j] = val[j]
                STAT["iters"] += 1; STAT["seqs"] += cnt; STAT["rounds"] += rounds
                ip += nx; op += e
                continue
            elif MEDIUM:
                # ---- medium step: one sequence whose literal and/or match length has ONE extension byte (< 255)
                tok = int(vb[0]); ll = tok >> 4; ml = tok & 15; pos = 1; okm = True
                if ll == 15:
                    x = int(vb[1])
                    if x == 255: okm = False
                    ll += x; pos = 2
                if okm and not (ip + pos + ll + 8 <= n and op + ll + 12 <= out_cap):
                    okm = False
                if okm:
                    q = ip + pos + ll
                    off = int(src[q]) | (int(src[q + 1]) << 8)
                    used = pos + ll + 2
                    if ml == 15:
                        x = int(src[q + 2])
                        if x == 255: okm = False
                        ml += x; used += 1
                    ml += 4
                    if okm and (off == 0 or off > op + ll or not (off >= 32 or off >= ml) or op + ll + ml + 12 > out_cap):
                        okm = False
                if okm:
                    i = 0
                    while i < ll:
                        for j in range(32):
                            if i + j < ll: out[op + i + j] = src[ip + pos + i + j]
                        i += 32
                    base = 0
                    while base < ml:
                        vals = [(out[op + ll + base + j - off] if base + j < ml else 0) for j in range(32)]
                        for j in range(32):
                            if base + j < ml: out[op + ll + base + j] = vals[j]
                        base += 32
                    ip += used; op += ll + ml
                    STAT["medium"] = STAT.get("medium", 0) + 1
                    continue
        STAT["slow"] += 1
        token = int(src[ip]); ip += 1
        ll = token >> 4
        if ll == 15:
            if ip >= n:
                return ("NONE", ip, out)
            while True:
                v = int(src[ip]); ip += 1
                ll = (ll + v) & 0xffffffff
                if not (v == 255 and ip < n - 15):
                    break
        if ll >= 0x80000000:
            return ("NONE", ip, out)
        lit_end = ip + ll
        lit_out_limit = op + ll
        if lit_out_limit > fast_output_limit - 4 or lit_end > n - 8:
            if lit_out_limit > out_cap:
                return ("LAST_LITERAL_OUTSIDE", ip, out)
            if lit_end != n:
                return ("ALL_INPUT_CONSUMED", ip, out)
            out[op:op + ll] = src[ip:ip + ll]
            op += ll
            break
        out[op:op + ll] = src[ip:ip + ll]
        op = lit_out_limit; ip = lit_end
        offset = int(src[ip]) | (int(src[ip + 1]) << 8); ip += 2
        if offset > op or offset == 0:
            return ("OFFSET_OUTSIDE", ip, out)
        ml = token & 15
        if ml == 15:
            while True:
                if ip > n - 5:
                    return ("NONE", ip, out)
                v = int(src[ip]); ip += 1
                ml = (ml + v) & 0xffffffff
                if v != 255:
                    break
        ml = (ml + 4) & 0xffffffff
        if ml >= 0x80000000:
            return ("NONE", ip, out)
        mol = op + ml
        if mol > fast_output_limit - 4 and mol > out_cap - 5:
            return ("LAST5_LITERALS", ip, out)
        for i in range(ml):
            out[op + i] = out[op + i - offset]
        op = mol
    return (0, op, out)


REASON = {1: "INPUT_EMPTY", 2: "LAST_LITERAL_OUTSIDE", 3: "ALL_INPUT_CONSUMED", 4: "OFFSET_OUTSIDE", 5: "LAST5_LITERALS", 6: "ZERO_CAP", 0: "NONE"}


def main(o=None, n_cases=40):
    import benchdata
    if o is None:
        from oracle.pyoracle import Oracle
        o = Oracle()
    full = os.path.join(benchdata.ROOT, "corpus", "silesia")
    blob = np.fromfile(os.path.join(benchdata.ROOT, "tests", "golden", "silesia_sample.bin"), dtype=np.uint8)
    rng = np.random.default_rng(5)
    starts = rng.integers(0, len(blob) - 16384, size=n_cases)
    cases = []
    for st in starts:
        sz = int(rng.choice([300, 2000, 8192, 16384]))
        raw = bytes(blob[st:st + sz])
        cases.append(raw)
    cases += [b"a" * 5000, bytes(range(256)) * 20, b"ab" * 3000, b"abc" * 2000 + b"xyz" * 700, (b"0123456789abcdefg" * 400)]
    checked = bad = 0
    for raw in cases:
        c = o.compress("lz4", raw)
        variants = [(c, len(raw)), (c, len(raw) + 100), (c, len(raw) - 1), (c, len(raw) + 11)]
        for _ in range(6):
            cc = bytearray(c)
            for _ in range(int(rng.integers(1, 3))):
                cc[int(rng.integers(0, len(cc)))] ^= int(rng.integers(1, 256))
            variants.append((bytes(cc), len(raw) + int(rng.integers(0, 64))))
        variants.append((c[:len(c) // 2], len(raw)))
        for comp, cap in variants:
            st, ln, out = decode(comp, cap)
            r, off, eout = o.decompress_raw("lz4", comp, cap)
            if r >= 0:
                exp = eout[:r].tobytes()
                eo = (0, r)
            else:
                exp = None
                eo = (REASON.get((-r) >> 8, str((-r) >> 8)), off)
            checked += 1
            if (st, ln) != eo or (exp is not None and bytes(out[:ln].astype(np.uint8)) != exp):
                bad += 1
                print("MISMATCH", (st, ln), eo, len(comp), cap)
    print("checked", checked, "bad", bad, STAT, "seq/iter", STAT["seqs"] / max(1, STAT["iters"]), "rounds/iter", STAT["rounds"] / max(1, STAT["iters"]))
    return checked, bad, dict(STAT)


if __name__ == "__main__":
    main()
