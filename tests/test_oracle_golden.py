"""CPU: the oracle (oracle/bz2_oracle.c) against the reference's own known-answer vectors and the
golden files minted from the unmodified reference (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import support as S
from golden.make_golden import stream_cases

G = S.GOLDEN


@pytest.mark.parametrize("i,level", [(1, 1), (2, 2), (3, 3)])
def test_reference_kat(i, level):
    """reference Makefile:58-66: bzip2 -1/-2/-3 of sample{1,2,3}.ref must equal sample{1,2,3}.bz2"""
    data = open(os.path.join(G, f"sample{i}.ref"), "rb").read()
    gold = open(os.path.join(G, f"sample{i}.bz2"), "rb").read()
    assert S.orc_compress(data, level) == gold


def test_golden_streams():
    gold = json.load(open(os.path.join(G, "streams.json")))
    seen = 0
    for name, data, level in stream_cases():
        g = gold[name]
        assert g["level"] == level and g["n"] == S.as_u8(data).size
        out = S.orc_compress(data, level)
        assert len(out) == g["out_len"], name
        assert hashlib.sha256(out).hexdigest() == g["sha256"], name
        seen += 1
    assert seen == len(gold)


def test_empty_and_single_byte_sizes():
    assert len(S.orc_compress(b"", 9)) == 14          # header + trailer only
    assert len(S.orc_compress(b"a", 9)) == 37


def test_crc_known_value():
    # CRC-32/BZIP2 check value of "123456789"
    assert S.oracle().orc_crc(S._p(S.as_u8(b"123456789")), 9) == 0xFC891918


def test_power_block_origptr_golden():
    """Exact powers: the oracle reports q and reproduces the reference's origPtr (golden) -- units with a single
    B* suffix and with several alike (oracle/tie_order.c replays the reference's tie order)."""
    gold = json.load(open(os.path.join(G, "origptr_powers.json")))
    for g in gold:
        blk = np.frombuffer(g["unit"].encode("latin-1") * g["q"], np.uint8)
        _, op, q = S.orc_bwt(blk)
        # q reported is the full multiplicity (e.g. "abab"^3 = "ab"^6)
        assert q >= g["q"] and q % g["q"] == 0
        assert op == g["orig_ptr"], g


def test_random_power_streams_golden():
    """240 seeded random (u, q): whole streams at -1 and -9 equal the reference's (sha256 minted by make_golden.py).
    The larger cases are thinned out here to keep the CPU suite short; the GPU suite runs all of them."""
    gold = json.load(open(os.path.join(G, "powers_random.json")))
    assert len(gold) >= 200
    checked = 0
    for k, g in enumerate(gold):
        n = g["p"] * g["q"]
        d = S.random_power_case(g["seed"], g["p"], g["alpha"], g["q"])
        levels = (1, 9) if n <= 100_000 else ((1,) if n <= 420_000 and k % 2 == 0 else ((9,) if k % 5 == 0 else ()))
        for level in levels:
            assert hashlib.sha256(S.orc_compress(d, level)).hexdigest() == g[f"sha_L{level}"], (g, level)
            checked += 1
        if "orig_ptr" in g and n <= 100_000:
            assert S.orc_bwt(S.orc_rle1_emit(d, 0, d.size)[0])[1] == g["orig_ptr"], g
    assert checked >= 250


def test_power_streams_match_reference_bytes():
    """Whole streams over exact-power blocks (constant data, "aab", 7-byte records, "abcabd" ...) are byte-identical."""
    gold = json.load(open(os.path.join(G, "streams.json")))
    for name, data, level in S.power_stream_cases():
        out = S.orc_compress(data, level)
        assert hashlib.sha256(out).hexdigest() == gold[name]["sha256"], name


def test_rle1_split_rules():
    """bzlib.c:211-315: chunks of <=255, block closes at the first chunk end at/after nblockMAX."""
    lvl = 1
    nmax = 100000 * lvl - 19
    d = np.concatenate([np.full(300, 7, np.uint8), (np.arange(nmax + 50, dtype=np.uint32) % 200 + 8).astype(np.uint8)])
    blocks = S.orc_split(d, lvl)
    assert len(blocks) == 2
    assert nmax <= blocks[0].nblock <= nmax + 4
    assert blocks[0].in_end == blocks[1].in_begin and blocks[1].in_end == d.size
    enc, inuse = S.orc_rle1_emit(d, 0, blocks[0].in_end)
    assert list(enc[:7]) == [7, 7, 7, 7, 251, 7, 7]      # 255-chunk then the remaining 45 -> 7,7,7,7,41
    assert inuse[251] == 1 and inuse[41] == 1
