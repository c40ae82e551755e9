"""CPU: the fast block split of run-free windows (csrc/stage1_rle.cu: k_has_run4 + k_chain_raw) stated as a small model and
checked against the oracle's split (orc_rle1_split, the restatement of bzlib.c:211-315).

When several engines work on one stream, a window's last block boundary is handed to the next engine before the window's
tiles have been scanned, if the window has no run of four or more equal bytes: then RLE1 leaves the bytes alone, a chunk (a
maximal run, here of 1-3 bytes) ends wherever the next byte differs, and a block closes at the first chunk end at or
after nblockMAX - 1 bytes from its start.  On the GPU the early result is compared with the full stage 1 of the same window
(a mismatch fails the job); this test pins the rule itself where no GPU is needed."""
import numpy as np

import support as S


def has_run4(a):
    """k_has_run4: is there a position q with a[q-3] == a[q-2] == a[q-1] == a[q]"""
    if a.size < 4:
        return False
    e = a[1:] == a[:-1]
    return bool(np.any(e[2:] & e[1:-1] & e[:-2]))


def chain_raw(a, nmax, blk_cap=1 << 30):
    """k_chain_raw: the loop of k_chain with the chunk-end flags computed from the raw bytes of a NON-final window"""
    W = a.size
    x = nb = 0
    while x < W and nb < blk_cap:
        target = x + nmax - 1
        if target >= W:
            break
        nxt = None
        for lane in range(8):
            e = target + lane
            if e + 1 < W and a[e + 1] != a[e]:
                nxt = e + 1
                break
        if nxt is None:
            break
        x, nb = nxt, nb + 1
    return x, nb


def short_run_data(rng, n, alpha):
    """bytes whose runs are 1-3 long: the inputs the fast split accepts"""
    # run values: a walk with non-zero steps modulo the alphabet, so neighbouring runs always differ
    steps = rng.integers(1, alpha, n).astype(np.int64) if alpha > 2 else np.ones(n, np.int64)
    vals = np.cumsum(steps) % alpha
    lens = rng.choice([1, 1, 1, 2, 2, 3], n)
    return np.repeat(vals, lens)[:n].astype(np.uint8)


def test_has_run4_model():
    assert not has_run4(np.frombuffer(b"aaabbbaaab", np.uint8))
    assert has_run4(np.frombuffer(b"abaaaab", np.uint8))
    assert has_run4(np.frombuffer(b"aaaa", np.uint8)) and not has_run4(np.frombuffer(b"aaa", np.uint8))


def test_fast_split_equals_oracle_split():
    rng = np.random.default_rng(31)
    nmax = 99_981
    checked = 0
    for it in range(60):
        alpha = int(rng.choice([2, 3, 4, 26, 256]))
        W = int(rng.integers(2 * nmax, 6 * nmax)) + int(rng.integers(0, 7))
        a = short_run_data(rng, W, alpha)
        if has_run4(a):
            # a merged run slipped through: cut it, the model under test is the chain, not the generator
            a = a.copy()
            e = a[1:] == a[:-1]
            bad = np.nonzero(e[2:] & e[1:-1] & e[:-2])[0] + 3
            a[bad] = ((a[bad].astype(np.int64) + 1) % max(alpha, 2)).astype(np.uint8)
            if has_run4(a):
                continue
        blocks = S.orc_split(a, 1)                 # the oracle closes its last block at the end of the data;
        x, nb = chain_raw(a, nmax)                 # a non-final window stops at the last COMPLETE block
        enc, _ = S.orc_rle1_emit(a, 0, a.size)
        assert np.array_equal(enc, a)              # RLE1 is the identity on such data
        assert nb == len(blocks) - 1 and x == blocks[-1].in_begin, (it, alpha, W, x, nb, len(blocks))
        for b in blocks[:-1]:
            assert nmax <= b.nblock <= nmax + 2    # runs of at most three bytes overshoot by at most two
        checked += 1
    assert checked >= 40


def test_fast_split_boundary_on_short_runs_straddling_the_limit():
    """the block limit falls inside a run of two or three: the boundary moves to the end of that run"""
    nmax = 99_981
    base = (np.arange(3 * nmax, dtype=np.uint32) % 251).astype(np.uint8)
    for runlen in (2, 3):
        for off in range(0, runlen + 1):
            a = base.copy()
            start = nmax - 1 - off
            a[start:start + runlen] = 252
            if has_run4(a):
                continue
            blocks = S.orc_split(a, 1)
            x, nb = chain_raw(a, nmax)
            assert nb == len(blocks) - 1 and x == blocks[-1].in_begin, (runlen, off)
            assert blocks[0].in_end == max(nmax, start + runlen) or blocks[0].in_end == nmax, (runlen, off, blocks[0].in_end)
