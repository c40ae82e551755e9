"""Regenerates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libbz2_ref.so).

Run in the authoring container only (needs /root/reference to have been compiled by
`make -C oracle ref`).  The GPU box has no reference tree; it uses the committed files.

  sample{1,2,3}.ref / .bz2   the reference's own known-answer vectors (Makefile:58-66), copied verbatim
  streams.json               sha256 + length of the reference's output for seeded synthetic inputs
  origptr_powers.json        the reference's origPtr on exact-power blocks u^q
  large_streams.json         (--large, ~1 min) sha256 of the reference's output for 120-200 MB inputs
  powers_random.json         240 seeded random (u, q): sha256 of the reference's stream at -1 and -9, and the
                             reference's origPtr when the input is a single block at -9
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import support as S  # noqa: E402


def stream_cases():
    """(name, bytes-like, level) -- every generator is seeded and lives in support.py / oracle/datagen.c"""
    rng = np.random.default_rng(12345)
    yield "empty", np.zeros(0, np.uint8), 9
    yield "one_byte", np.array([97], np.uint8), 9
    yield "text_200k_L1", S.gen_text(200_000), 1
    yield "text_200k_L9", S.gen_text(200_000), 9
    yield "text_2M_L9", S.gen_text(2_000_000), 9
    yield "text_1M_L3", S.gen_text(1_000_000, seed=99), 3
    yield "random_300k_L2", S.gen_random(300_000), 2
    yield "random_1M_L9", S.gen_random(1_000_000, seed=5), 9
    yield "period1000_1M_L9", S.gen_period1000(1_000_000), 9
    yield from S.power_stream_cases()
    yield "period1000_500k_L1", S.gen_period1000(500_000), 1
    yield "aab_1M_L9", S.gen_tile(1_000_000, b"aab"), 9
    yield "runs_4M_L9", S.gen_runs(4_000_000), 9
    yield "runs_1M_L1", S.gen_runs(1_000_000, seed=11), 1
    yield "mixed_3M_L5", S.gen_mixed(3_000_000, seg=1 << 18), 5
    yield "fb_2M_L9", np.full(2_000_000, 251, np.uint8), 9
    for k in (254, 255, 256, 257, 600):
        yield f"run{k}_L9", np.concatenate([np.full(k, 65, np.uint8), np.array([66, 67], np.uint8)]), 9
    yield "alpha2_50k_L9", rng.integers(0, 2, 50_000, dtype=np.uint8), 9
    yield "alpha256_70k_L1", rng.integers(0, 256, 70_000, dtype=np.uint8), 1
    # block-fill corner: block fills exactly one byte before the end (bzlib.c:276-308)
    yield "tailmerge_L1", (np.arange(99981 + 1, dtype=np.uint32) % 251).astype(np.uint8), 1
    yield "tailmerge2_L1", (np.arange(99981 + 2, dtype=np.uint32) % 251).astype(np.uint8), 1
    # long non-tandem repeats (stage 2, repeat passes)
    for name, d in S.long_repeat_cases():
        yield "rep_" + name + "_L9", d, 9


def large_cases():
    """Window-scale inputs: >= 2 full windows of ~111 blocks each at -9, ~1200 blocks at -1."""
    yield "text_200M_L9", S.gen_text(200_000_000), 9
    yield "c4_200M_L9", S.gen_c4(200_000_000, seg=64 << 20), 9
    yield "text_120M_L1", S.gen_text(120_000_000, seed=7), 1


def power_cases():
    units = [b"ab", b"ba", b"abc", b"aab", b"cab", b"abcabd", b"1234567", b"a", b"zyx", b"abab",
             b"\x00\x00\x00\x00\xfb", b"\xff\xff\xff\xff\xfb", b"aabb", b"acb", b"abcdcb", b"qqzzq"]
    qs = [2, 3, 8, 9, 10, 11, 12, 13, 100, 101, 1000, 1001, 1024, 1025, 1026, 1027, 1028, 2048, 2049, 5000]
    for u in units:
        for q in qs:
            if len(u) * q <= 60000:
                yield u, q
    yield b"abc", 33327
    yield b"ab", 449990
    yield b"abc", 299993
    yield b"abcabd", 149996
    yield b"cab", 299993


def random_power_params():
    """(seed, p, alpha, q).  Small: one block at both levels.  Divisor periods of nblockMAX(-1) = 99,981 =
    27*7*529: every full block at -1 is an exact power.  Large: one block of up to 899,981 bytes at -9."""
    import random
    rng = random.Random(20261018)
    out = []
    for k in range(130):
        p = rng.randint(2, 40)
        alpha = rng.choice([2, 3, 4, 6, 256])
        q = rng.choice([2, 3, 9, 10, 11, 100, 513, 1024, 1025, 1026, 1027, 1028, 2049, rng.randint(2, 4000)])
        q = max(2, min(q, 99_000 // p))
        out.append((1000 + k, p, alpha, q))
    for k in range(80):
        p = rng.choice([3, 7, 9, 21, 23, 27, 63, 69, 161, 189, 207, 483, 529, 621, 1587, 3703])
        alpha = rng.choice([3, 4, 5, 26, 256])
        out.append((2000 + k, p, alpha, rng.randint(2 * 99_981 // p, 4 * 99_981 // p) + 1))
    for k in range(30):
        p = rng.choice([rng.randint(41, 3000), rng.randint(3000, 150_000), rng.randint(150_000, 449_990)])
        alpha = rng.choice([2, 4, 256])
        qmax = 899_981 // p
        out.append((3000 + k, p, alpha, max(2, rng.choice([2, 3, qmax, rng.randint(2, qmax)]))))
    return out


def main():
    streams = {}
    for name, data, level in stream_cases():
        out = S.ref_compress(data, level)
        streams[name] = {"level": level, "n": int(S.as_u8(data).size), "out_len": len(out),
                         "sha256": hashlib.sha256(out).hexdigest()}
        print(name, len(out))
    json.dump(streams, open(os.path.join(HERE, "streams.json"), "w"), indent=1, sort_keys=True)
    powers = []
    for u, q in power_cases():
        blk = np.frombuffer(u * q, np.uint8)
        _, op = S.ref_bwt(blk)
        powers.append({"unit": u.decode("latin-1"), "q": q, "orig_ptr": int(op)})
    json.dump(powers, open(os.path.join(HERE, "origptr_powers.json"), "w"), indent=0)
    print(len(powers), "power cases")
    rnd = []
    for seed, p, alpha, q in random_power_params():
        d = S.random_power_case(seed, p, alpha, q)
        rec = {"seed": seed, "p": p, "alpha": alpha, "q": q}
        for level in (1, 9):
            rec[f"sha_L{level}"] = hashlib.sha256(S.ref_compress(d, level)).hexdigest()
        recs = S.ref_trace(d, 9)[0]
        if len(recs) == 1:                                  # one block at -9: the reference's origPtr for it
            rec["orig_ptr"] = int(recs[0].orig_ptr)
        rnd.append(rec)
    json.dump(rnd, open(os.path.join(HERE, "powers_random.json"), "w"), indent=0)
    print(len(rnd), "random power cases,", sum("orig_ptr" in r for r in rnd), "with origPtr")
    if "--large" in sys.argv:
        large = {}
        for name, data, level in large_cases():
            out = S.ref_compress(data, level)
            large[name] = {"level": level, "n": int(data.size), "out_len": len(out), "sha256": hashlib.sha256(out).hexdigest()}
            print(name, len(out))
        json.dump(large, open(os.path.join(HERE, "large_streams.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
