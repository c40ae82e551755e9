"""CPU: differential check of the oracle against the unmodified reference (oracle/_ref), when present.

The reference tree only exists in the authoring container; on other boxes these tests skip and the
committed golden files (test_oracle_golden.py) carry the pin."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def _same_or_origptr_only(data, level):
    r = S.ref_compress(data, level)
    o = S.orc_compress(data, level)
    if r == o:
        return True
    recs, _, _ = S.ref_trace(data, level)
    return S.orc_compress(data, level, force_orig_ptr=[x.orig_ptr for x in recs]) == r


def test_fuzz_small():
    rng = np.random.default_rng(2024)
    for it in range(250):
        n = int(rng.integers(0, 4000))
        alpha = int(rng.integers(1, 257))
        mode = it % 4
        if mode == 0:
            d = rng.integers(0, alpha, n, dtype=np.uint8)
        elif mode == 1:
            d = np.repeat(rng.integers(0, alpha, n // 7 + 1, dtype=np.uint8), rng.integers(1, 300, n // 7 + 1))[:n].astype(np.uint8)
        elif mode == 2:
            d = np.resize(rng.integers(0, alpha, int(rng.integers(1, 40)), dtype=np.uint8), n)
        else:
            d = S.gen_text(n, seed=it + 1)
        assert _same_or_origptr_only(d, int(rng.integers(1, 10))), (it, mode, n)


@pytest.mark.parametrize("gen,n,level", [
    ("text", 1_200_000, 1), ("random", 700_000, 2), ("runs", 3_000_000, 1), ("p1000", 450_000, 1), ("mixed", 900_000, 2),
])
def test_multiblock(gen, n, level):
    d = {"text": S.gen_text, "random": S.gen_random, "runs": S.gen_runs, "p1000": S.gen_period1000,
         "mixed": lambda k: S.gen_mixed(k, seg=100_000)}[gen](n)
    assert S.ref_compress(d, level) == S.orc_compress(d, level)


def test_stage_functions_match_reference():
    d = S.gen_text(150_000)
    recs, blk, _ = S.ref_trace(d, 1, want_block=0)
    blocks = S.orc_split(d, 1)
    assert [b.nblock for b in blocks] == [r.nblock for r in recs]
    assert [b.crc for b in blocks] == [r.block_crc for r in recs]
    enc, inuse = S.orc_rle1_emit(d, blocks[0].in_begin, blocks[0].in_end)
    assert np.array_equal(enc, blk[: blocks[0].nblock])
    rb, rop = S.ref_bwt(enc)
    ob, oop, q = S.orc_bwt(enc)
    assert q == 1 and rop == oop and np.array_equal(rb, ob)
    m, f, nu = S.orc_mtf(ob, inuse)
    assert len(m) == recs[0].n_mtf and nu == recs[0].n_in_use


def test_tail_merge_corner():
    """bzlib.c:276-308: a lone last byte joins a block that has just filled (BuffToBuff semantics)."""
    nmax = 99981
    d = (np.arange(nmax + 1, dtype=np.uint32) % 251).astype(np.uint8)
    assert len(S.ref_trace(d, 1)[0]) == 1
    assert S.ref_compress(d, 1) == S.orc_compress(d, 1)
    assert len(S.orc_split(d, 1, tail_merge=0)) == 2


def test_power_offset_rule_against_reference():
    """Exact powers of units with one B* suffix: lo + orc_power_offset == the reference's origPtr."""
    import random
    rng = random.Random(77)
    done = 0
    while done < 120:
        k = rng.randint(2, 9)
        vals = sorted(rng.sample(range(256), k))
        up = []
        for v in vals:
            up += [v] * rng.choice([1, 1, 2, 5])
        down = []
        for v in reversed(vals[1:-1]):
            if rng.random() < 0.6:
                down += [v] * rng.choice([1, 2])
        u = bytes(up + down)
        r = rng.randrange(len(u))
        u = u[r:] + u[:r]
        p = len(u)
        if any(p % d == 0 and u == u[:d] * (p // d) for d in range(1, p)):
            continue
        q = rng.choice([2, 5, 9, 10, 11, 64, 65, 1024, 1025, 1026, 1027, 1028, 3000, rng.randint(2, 2000)])
        if p * q > 200_000:
            continue
        blk = np.frombuffer(u * q, np.uint8)
        off = S.orc_power_offset(blk, q)
        if off < 0:
            continue                                    # several B* suffixes: outside the rule
        _, lo, qq = S.orc_bwt(blk)
        _, op = S.ref_bwt(blk)
        assert qq == q and lo + off == op, (u, q, lo, off, op)
        done += 1
