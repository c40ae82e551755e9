"""CPU: differential check of the oracle against the unmodified reference (oracle/_ref), when present.

The reference tree only exists in the authoring container; on other boxes these tests skip and the
committed golden files (test_oracle_golden.py) carry the pin."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


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
        level = int(rng.integers(1, 10))
        assert S.ref_compress(d, level) == S.orc_compress(d, level), (it, mode, n)


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


def test_tie_order_against_reference():
    """Exact powers u^q, units with one or several B* suffixes: the oracle's origPtr (canonical rank + the replayed tie
    order, oracle/tie_order.c) equals the reference's on random (u, q), including the 1024-chunk and budget regimes."""
    import random
    rng = random.Random(77)
    for it in range(500):
        p = rng.choice([2, 3, 4, 5, 6, 7, 9, 12, 21, 23, 27, 50, 100, 333, 1000, rng.randint(2, 5000)])
        alpha = rng.choice([2, 3, 4, 8, 256])
        u = bytes(rng.randrange(alpha) for _ in range(p))
        q = rng.choice([2, 3, 5, 8, 9, 10, 11, 50, 100, 513, 1000, 1024, 1025, 1026, 1027, 2000, 3000, 5000])
        q = max(2, min(q, 250_000 // p))
        blk = np.frombuffer(u * q, np.uint8)
        ob, oop, qq = S.orc_bwt(blk)
        rb, rop = S.ref_bwt(blk)
        assert qq >= q and qq % q == 0
        assert np.array_equal(ob, rb) and oop == rop, (u[:32], p, q, oop, rop)


@pytest.mark.parametrize("unit,n,level", [(b"ab\ncd\n.", 420_000, 1), (b"ab\ncd\n.", 420_000, 9), (b"abcabd", 300_000, 1),
                                          (b"aabb", 250_000, 2), (b"abcdcb", 99_981 * 2, 1)])
def test_periodic_records_whole_stream(unit, n, level):
    """fixed-width records whose blocks are exact powers with several B* suffixes per unit."""
    d = S.gen_tile(n, unit)
    assert S.ref_compress(d, level) == S.orc_compress(d, level)
