"""CPU: a small numpy/Python model of stage 2's refinement rounds WITH the repeat passes (stage2_bwt.cu, 2e),
checked against a brute-force rotation sort.

The CUDA kernels are checked bit for bit against the oracle on the GPU (test_gpu_parity.py).  This file checks the
*rules* themselves, independently of CUDA, on thousands of small inputs built to have long non-tandem repeats:

  segment chains    S is chained when the successors of its members all lie in one segment; after j chained
                    steps S may be sorted by the rank at offset j + d            (k_rep_small / k_rep_warp / k_rep_dist)
  dominant offset   eq[x] = (rank[x] == rank[x + delta]); J[x] = length of that run; a segment whose members are a
                    progression of step delta is sorted by the rank at offset min J, and groups of equal keys
                    that are again such progressions take further keys in the same visit   (k_rep_eq, k_refine_small)
  relaxed tandem    a progression of step delta whose neighbour pairs all (but the top one) repeat for >= delta
                    symbols is monotone; the top pair decides the direction                 (k_resolve_periodic)

Every round must keep the invariant the kernels rely on: the members of a segment agree on at least `depth` symbols,
and the group-start ranks are consistent with the true rotation order.  Test infrastructure only."""
import numpy as np
import pytest


def next_zero_distance(flag):
    """j[x] = distance from x to the next position (cyclically, at or after x) whose flag is 0; all 0 if there is none."""
    n = len(flag)
    j = np.zeros(n, np.int64)
    zeros = np.flatnonzero(flag == 0)
    if len(zeros) == 0:
        return j
    nxt = zeros[0] + n
    for x in range(n - 1, -1, -1):
        if flag[x] == 0:
            nxt = x
        j[x] = nxt - x
    return j


def refine_with_repeat_passes(T, use_chain=True, use_delta=True, use_tandem=True, max_rounds=64):
    """Prefix doubling over the cyclic rotations of T.  Returns (rank, rounds): rank[i] = start of i's group in the
    sorted order; rotations left in one group at the end are equal."""
    T = np.asarray(T, np.int64)
    n = len(T)
    order = np.argsort(T, kind="stable")
    rank = np.empty(n, np.int64)
    sa = order.copy()
    start = 0
    for k in range(1, n + 1):
        if k == n or T[sa[k]] != T[sa[start]]:
            rank[sa[start:k]] = start
            start = k
    d, rounds = 1, 0
    while d < n and rounds < max_rounds:
        segs = []
        s = 0
        for k in range(1, n + 1):
            if k == n or rank[sa[k]] != rank[sa[s]]:
                if k - s >= 2:
                    segs.append((s, k))
                s = k
        if not segs:
            break
        rounds += 1
        old = rank.copy()
        # --- repeat passes, from the ranks the round starts with
        cflag = np.zeros(n, np.int64)
        votes = {}
        for s, e in segs:
            mem = sa[s:e]
            succ = old[(mem + 1) % n]
            if use_chain and np.all(succ == succ[0]):
                cflag[mem] = 1
            dl = abs(int(mem[0]) - int(mem[1]))
            votes[dl] = votes.get(dl, 0) + (e - s)
        jd = next_zero_distance(cflag)
        delta = max(votes, key=lambda k: (votes[k], k)) if (use_delta and votes) else 0
        if delta and delta < n:
            eq = (old == old[(np.arange(n) + delta) % n]).astype(np.int64)
            jq = next_zero_distance(eq)
        else:
            delta, jq = 0, np.zeros(n, np.int64)
        new = old.copy()
        for s, e in segs:
            mem = np.sort(sa[s:e])
            m = len(mem)
            # relaxed tandem rule
            if use_tandem and delta and m >= 3 and np.all(np.diff(mem) == delta):
                js = jq[mem[:-1]]
                if np.all(js[:-1] >= delta):
                    jtop = js[-1]
                    if jtop < delta:
                        ra, rb = old[(mem[-2] + jtop) % n], old[(mem[-1] + jtop) % n]
                        decided = ra != rb
                        asc = ra < rb
                    else:
                        xr = old[(mem[-1] + delta) % n]
                        decided = not (s <= xr < e)
                        asc = xr >= e
                    if decided:
                        seq = mem if asc else mem[::-1]
                        sa[s:e] = seq
                        new[seq] = s + np.arange(m)
                        continue
            off = d + (jd[mem[0]] if use_chain else 0)
            assert not use_chain or np.all(jd[mem] == jd[mem[0]])        # the chain length is a property of the segment
            out = _sort_group(mem, int(off), True, delta, jq, old, n)
            pos = s
            for pc in out:
                sa[pos:pos + len(pc)] = pc
                new[pc] = pos
                pos += len(pc)
            assert pos == e
        rank = new
        d *= 2
    return rank, rounds


def _sort_group(mem, off, first, delta, jq, old, n):
    """One visit of a segment: the first key at offset `off` (raised to the shortest pairwise repeat if the members are
    a progression of step delta); groups of equal keys that are again such progressions, with a longer shortest
    repeat, take further keys.  Returns the pieces in sorted order; members of a piece had equal keys throughout."""
    g = np.array(sorted(int(x) for x in mem))
    gs = set(g.tolist())
    o = off if first else None
    if delta:
        has = np.array([(x + delta < n) and ((x + delta) in gs) for x in g])     # not cyclic, as in the kernel
        if len(g) >= 2 and has.sum() + 1 == len(g):
            jm = int(jq[g[has]].min())
            if first:
                o = max(off, jm)
            elif jm > off:
                o = jm
    if o is None:
        return [list(g)]
    keys = old[(g + o) % n]
    idx = np.argsort(keys, kind="stable")
    g, keys = g[idx], keys[idx]
    res, a = [], 0
    for k in range(1, len(g) + 1):
        if k == len(g) or keys[k] != keys[a]:
            piece = list(g[a:k])
            if 2 <= len(piece) < len(g):
                res.extend(_sort_group(piece, o, False, delta, jq, old, n))
            else:
                res.append(piece)
            a = k
    return res


def brute_ranks(T):
    """group-start rank of every rotation under the true cyclic order (equal rotations share a group)"""
    T = list(T)
    n = len(T)
    rots = sorted(range(n), key=lambda i: T[i:] + T[:i])
    rank = [0] * n
    s = 0
    for k in range(1, n + 1):
        if k == n or (T[rots[k]:] + T[:rots[k]]) != (T[rots[s]:] + T[:rots[s]]):
            for i in rots[s:k]:
                rank[i] = s
            s = k
    return np.array(rank)


def _cases(rng, count):
    for it in range(count):
        mode = it % 6
        alpha = int(rng.integers(2, 6))
        if mode == 0:                                    # x . noise . x . x[shifted]
            x = rng.integers(0, alpha, int(rng.integers(8, 60)))
            y = rng.integers(0, alpha, int(rng.integers(1, 9)))
            T = np.concatenate([x, y, x, x[len(x) // 3:]])
        elif mode == 1:                                  # tiled unit cut off mid-period (several copies per block)
            u = rng.integers(0, alpha, int(rng.integers(3, 40)))
            T = np.resize(u, int(rng.integers(len(u) + 1, 5 * len(u) + 7)))
        elif mode == 2:                                  # exact power
            u = rng.integers(0, alpha, int(rng.integers(1, 12)))
            T = np.tile(u, int(rng.integers(2, 9)))
        elif mode == 3:                                  # copy with one edit
            x = rng.integers(0, alpha, int(rng.integers(10, 80)))
            z = x.copy()
            z[int(rng.integers(0, len(z)))] = alpha
            T = np.concatenate([x, z, x[: len(x) // 2]])
        elif mode == 4:                                  # three different repeat offsets in one block
            a, b, c = (rng.integers(0, alpha, int(rng.integers(5, 25))) for _ in range(3))
            T = np.concatenate([a, b, a, c, b, c, a])
        else:
            T = rng.integers(0, alpha, int(rng.integers(2, 150)))
        yield T


@pytest.mark.parametrize("flags", [(True, True, True), (True, False, False), (False, True, False), (False, True, True), (False, False, False)])
def test_repeat_rules_give_the_true_rotation_order(flags):
    rng = np.random.default_rng(31)
    for T in _cases(rng, 600):
        got, _ = refine_with_repeat_passes(T, *flags)
        assert np.array_equal(got, brute_ranks(T)), (flags, T.tolist())


def test_repeat_passes_cut_the_rounds_on_long_repeats():
    rng = np.random.default_rng(7)
    x = rng.integers(0, 256, 700)
    T = np.resize(x, 2000)                               # ~3 copies, cut mid-period: every rotation has a twin 700 further on
    plain, r0 = refine_with_repeat_passes(T, False, False, False)
    fast, r1 = refine_with_repeat_passes(T, True, True, True)
    assert np.array_equal(plain, fast) and np.array_equal(fast, brute_ranks(T))
    assert r0 >= 9 and r1 <= 4, (r0, r1)
