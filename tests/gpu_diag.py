"""Stage-by-stage GPU-vs-oracle diagnosis (run on a GPU box; prints the first mismatch per stage).

    python tests/gpu_diag.py [case ...]

Not a pytest file: it is the tool used to localise a parity failure to one kernel family.
"""
import sys
import os
import time
import traceback
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import support as S  # noqa: E402
import bzip2_b200 as B  # noqa: E402


def first_diff(a, b):
    n = min(len(a), len(b))
    a = np.asarray(a[:n]); b = np.asarray(b[:n])
    d = np.nonzero(a != b)[0]
    if d.size:
        return int(d[0])
    return None if len(a) == len(b) else n


def diag(name, data, level, eng):
    data = S.as_u8(data)
    t0 = time.time()
    try:
        got = eng.compress(data)
    except Exception as ex:  # noqa: BLE001
        print(f"[{name}] L{level} n={data.size}: EXCEPTION {ex}")
        return False
    t1 = time.time()
    exp = S.orc_compress(data, level)
    st = eng.stats
    ok = got == exp
    print(f"[{name}] L{level} n={data.size} out={len(got)} exp={len(exp)} blocks={st.n_blocks} "
          f"ms={st.ms_total:.2f} (s1 {st.ms_s1:.2f} s2 {st.ms_s2:.2f} s3 {st.ms_s3:.2f} s4 {st.ms_s4:.2f}) wall={t1-t0:.3f}s "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
    if ok:
        return True
    fd = first_diff(np.frombuffer(got, np.uint8), np.frombuffer(exp, np.uint8))
    print(f"    first differing output byte: {fd}")
    # only the last window is inspectable; use inputs that fit one window for diagnosis
    blocks = S.orc_split(data, level)
    X = eng.fetch("X", np.uint32)
    P = eng.fetch("P", np.uint32)
    nb = len(X) - 1
    print(f"    blocks: gpu {nb} oracle {len(blocks)}")
    exp_X = np.concatenate([[0], np.cumsum([b.nblock for b in blocks])]).astype(np.uint64)
    exp_P = np.array([b.in_begin for b in blocks] + [data.size], dtype=np.uint64)
    if nb != len(blocks) or first_diff(X, exp_X) is not None:
        print(f"    S1 X mismatch: gpu {X[:6]}..{X[-3:]} exp {exp_X[:6]}..{exp_X[-3:]}")
        return False
    if first_diff(P, exp_P) is not None:
        print(f"    S1 P mismatch at {first_diff(P, exp_P)}: gpu {P[:6]} exp {exp_P[:6]}")
        return False
    crc = eng.fetch("crc", np.uint32)
    exp_crc = np.array([b.crc for b in blocks], np.uint32)
    if first_diff(crc, exp_crc) is not None:
        k = first_diff(crc, exp_crc)
        print(f"    S1 CRC mismatch block {k}: gpu {crc[k]:08x} exp {exp_crc[k]:08x}")
    enc = eng.fetch("enc", np.uint8)
    bwt = eng.fetch("bwt", np.uint8)
    mtfv = eng.fetch("mtfv", np.uint16)
    nmtf = eng.fetch("nmtf", np.uint32)
    op = eng.fetch("origptr", np.uint32)
    inuse = eng.fetch("inuse", np.uint8).reshape(nb, 256)
    freq = eng.fetch("mtffreq", np.int32).reshape(nb, 258)
    pq = eng.fetch("power_q", np.uint32)
    bits = eng.fetch("bits", np.uint64)
    bad = 0
    for b, blk in enumerate(blocks):
        x0, x1 = int(X[b]), int(X[b + 1])
        e_exp, iu_exp = S.orc_rle1_emit(data, blk.in_begin, blk.in_end)
        d = first_diff(enc[x0:x1], e_exp)
        if d is not None:
            print(f"    S1 enc mismatch block {b} at {d}: gpu {enc[x0+d-2:x0+d+6]} exp {e_exp[max(0,d-2):d+6]}")
            bad += 1; break
        if first_diff(inuse[b], iu_exp) is not None:
            print(f"    S2 inuse mismatch block {b}")
            bad += 1
        bw_exp, op_exp, q = S.orc_bwt(e_exp)
        d = first_diff(bwt[x0:x1], bw_exp)
        if d is not None:
            print(f"    S2 BWT mismatch block {b} (n={x1-x0}) at {d}; power_q gpu={pq[b]} oracle q={q}")
            sa = eng.fetch("sa", np.uint32)[x0:x1]
            print(f"       sa is permutation: {np.array_equal(np.sort(sa), np.arange(x1-x0))}")
            bad += 1
            if bad > 3:
                break
            continue
        if op[b] != op_exp:
            print(f"    S2 origPtr mismatch block {b}: gpu {op[b]} exp {op_exp} (q={q}, power_q={pq[b]})")
            bad += 1
        m_exp, f_exp, nu = S.orc_mtf(bw_exp, iu_exp)
        if nmtf[b] != len(m_exp):
            print(f"    S3 nMTF mismatch block {b}: gpu {nmtf[b]} exp {len(m_exp)}")
            bad += 1
        mb = x0 + b
        d = first_diff(mtfv[mb:mb + len(m_exp)], m_exp)
        if d is not None:
            print(f"    S3 mtfv mismatch block {b} at {d}: gpu {mtfv[mb+d-3:mb+d+5]} exp {m_exp[max(0,d-3):d+5]}")
            z = eng.fetch("z", np.uint8)[x0:x1]
            print(f"       z head {z[:24]}")
            bad += 1
            if bad > 3:
                break
            continue
        if first_diff(freq[b], f_exp) is not None:
            k = first_diff(freq[b], f_exp)
            print(f"    S3 mtfFreq mismatch block {b} sym {k}: gpu {freq[b][k]} exp {f_exp[k]}")
            bad += 1
        # S4: coded size
        import ctypes as C
        tmp = np.zeros(len(m_exp) * 3 + 70000, np.uint8)
        bp = C.c_uint64(0)
        S.oracle().orc_send_mtf(S._p(m_exp, C.c_uint16), len(m_exp), S._p(iu_exp), S._p(f_exp, C.c_int32), S._p(tmp), C.byref(bp))
        exp_bits = bp.value + 48 + 32 + 1 + 24
        if bits[b] != exp_bits:
            print(f"    S4 coded size mismatch block {b}: gpu {bits[b]} exp {exp_bits}")
            bad += 1
            if bad > 3:
                break
    if bad == 0:
        print("    all per-block stage outputs match; mismatch is in S4 bit emission / assembly")
        bo = eng.fetch("bitoff", np.uint64)
        print(f"       bitoff {bo[:4]} .. {bo[-2:]}; first diff byte {fd} = bit {fd*8 if fd is not None else None}")
    return False


def cases():
    rng = np.random.default_rng(7)
    yield "empty", b"", 9
    yield "one", b"a", 9
    yield "tiny", b"hello hello hello world", 9
    yield "aaaa", b"a" * 10, 9
    yield "runs", np.repeat(rng.integers(0, 5, 400, dtype=np.uint8), rng.integers(1, 600, 400)), 9
    yield "text20k", S.gen_text(20000), 9
    yield "rand20k", S.gen_random(20000), 9
    yield "text300k_L1", S.gen_text(300000), 1
    yield "rand250k_L1", S.gen_random(250000), 1
    yield "text2M", S.gen_text(2_000_000), 9
    yield "rand2M", S.gen_random(2_000_000), 9
    yield "p1000_1M", S.gen_period1000(1_000_000), 9
    yield "aab_1M", S.gen_tile(1_000_000, b"aab"), 9
    yield "runs3M", S.gen_runs(3_000_000), 9
    yield "mixed3M_L3", S.gen_mixed(3_000_000, seg=1 << 18), 3
    yield "fb_3M", np.full(3_000_000, 251, np.uint8), 9


def main():
    want = set(sys.argv[1:])
    engines = {}
    nok = nbad = 0
    for name, data, level in cases():
        if want and name not in want:
            continue
        if level not in engines:
            engines[level] = B.Engine(level=level)
        try:
            ok = diag(name, data, level, engines[level])
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            ok = False
        nok += ok; nbad += (not ok)
    print(f"diag: {nok} ok, {nbad} bad")
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
