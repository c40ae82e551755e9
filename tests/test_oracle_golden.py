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


def test_power_block_bwt_and_period():
    """Exact powers: the oracle's BWT bytes are canonical and it reports q; the reference's origPtr
    (golden) always lies inside the tie group [lo, lo+q)."""
    gold = json.load(open(os.path.join(G, "origptr_powers.json")))
    checked = 0
    for g in gold:
        if len(g["unit"]) * g["q"] > 70000:
            continue
        blk = np.frombuffer(g["unit"].encode("latin-1") * g["q"], np.uint8)
        _, lo, q = S.orc_bwt(blk)
        # q reported is the full multiplicity (e.g. "abab"^3 = "ab"^6)
        assert q >= g["q"] and q % g["q"] == 0
        assert lo <= g["orig_ptr"] < lo + q, g
        off = S.orc_power_offset(blk, q)
        if off >= 0:                       # unit with a single B* suffix: the reference's choice is reproduced
            assert lo + off == g["orig_ptr"], g
            checked += 1
    assert checked >= 100


def test_power_streams_match_reference_bytes():
    """Whole streams over exact-power blocks with single-B* units (constant data, "aab") are byte-identical."""
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
