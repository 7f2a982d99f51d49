"""Lane-level model of search_warp.cuh::mm_push_warp against the sequential MinMaxHeap::push (search_core.cuh::mm_push,
min_max_heap crate semantics): the warp-cooperative bubble-up must leave the heap array in exactly
the same state.  Pure Python (the kernel itself is covered on the GPU by the retry-lane parity tests)."""
import random


def on_min_level(i):
    return ((i + 1).bit_length() - 1) % 2 == 0


def push_seq(h, e):
    """mm_push: parent step, then grandparent climbing."""
    i = len(h)
    h.append(None)
    if i > 0:
        p = (i - 1) >> 1
        if on_min_level(i):
            if e > h[p]:
                h[i] = h[p]; i = p; climb_max = True
            else:
                climb_max = False
        else:
            if e < h[p]:
                h[i] = h[p]; i = p; climb_max = False
            else:
                climb_max = True
    else:
        climb_max = not on_min_level(i)
    while i >= 3:
        gp = (((i - 1) >> 1) - 1) >> 1
        if (e > h[gp]) if climb_max else (e < h[gp]):
            h[i] = h[gp]; i = gp
        else:
            break
    h[i] = e


def push_warp(h, e):
    """mm_push_warp, lane by lane: all loads first, one ballot, then the writes."""
    i = len(h)
    h.append(None)
    if i == 0:
        h[0] = e
        return
    p = (i - 1) >> 1
    min_level = on_min_level(i)
    lanes = []
    for lane in range(32):
        in_a, in_b = 1 <= lane <= 15, 16 <= lane <= 30
        lvl = lane if in_a else (lane - 15 if in_b else 0)
        base1 = (p if in_b else i) + 1
        anc1 = base1 >> (2 * lvl) if lvl else 0
        valid = lvl != 0 and anc1 >= 1
        v = h[p] if lane == 0 else (h[anc1 - 1] if valid else e)
        lanes.append((in_a, in_b, lvl, base1, valid, v))
    pe = lanes[0][5]
    moved = (e > pe) if min_level else (e < pe)
    climb_max = min_level == moved
    ball = 0
    for lane, (in_a, in_b, lvl, base1, valid, v) in enumerate(lanes):
        if valid and ((e > v) if climb_max else (e < v)):
            ball |= 1 << lane
    chain = ((ball >> 16) if moved else (ball >> 1)) & 0x7FFF
    t = 0
    while (chain >> t) & 1:
        t += 1
    cur = p if moved else i
    writes = []
    for lane, (in_a, in_b, lvl, base1, valid, v) in enumerate(lanes):
        if lane == 0:
            if moved:
                writes.append((i, pe))
            fin = cur if t == 0 else ((cur + 1) >> (2 * t)) - 1
            writes.append((fin, e))
        elif valid and (in_b if moved else in_a) and lvl <= t:
            below = cur if lvl == 1 else (base1 >> (2 * (lvl - 1))) - 1
            writes.append((below, v))
    assert len({w[0] for w in writes}) == len(writes), "lanes must write distinct slots"
    for pos, val in writes:
        h[pos] = val


def pop_max(h):
    """enough of MinMaxHeap::pop_max to keep the model heaps realistic (both copies get the same treatment)"""
    if len(h) <= 2:
        return h.pop() if h else None
    m = 1 if len(h) == 2 or h[1] > h[2] else 2
    h[m], h[-1] = h[-1], h[m]
    return h.pop()


def test_warp_push_matches_sequential_push():
    rng = random.Random(7)
    for trial in range(60):
        a, b = [], []
        span = rng.choice([3, 10, 1000, 10 ** 6])  # few distinct keys -> many ties
        for step in range(rng.choice([5, 40, 700, 5000])):
            if a and rng.random() < 0.2:
                # a plain swap-pop breaks the heap property in both copies alike; the push code only ever compares
                # along ancestor chains, so equality of the two arrays is still the right check
                pop_max(a); pop_max(b)
                continue
            # best-first search mostly pushes near-maximal keys
            e = (-(rng.randrange(span) if rng.random() < 0.7 else 0), step)
            push_seq(a, e)
            push_warp(b, e)
            assert a == b, (trial, step)


class SparseHeap:
    """list-like with a default content, so that multi-million-entry heaps cost nothing"""
    def __init__(self, n):
        self.n, self.d = n, {}
    def __len__(self):
        return self.n
    def append(self, v):
        self.d[self.n] = v
        self.n += 1
    def __getitem__(self, k):
        assert 0 <= k < self.n
        return self.d.get(k, (k % 97, k))
    def __setitem__(self, k, v):
        assert 0 <= k < self.n
        self.d[k] = v


def test_deep_heap_indices():
    # indices near the reference's STACK_LIMIT (2e6) and EDIT_TREE_LIMIT (1e7): chains of up to 12 grandparent levels
    for n in (2_000_000, 2_000_001, 10_000_031, (1 << 24) - 2):
        a, b = SparseHeap(n), SparseHeap(n)
        for e in ((1000, -1), (-1000, -2), (48, -3), (96, -4)):
            push_seq(a, e)
            push_warp(b, e)
            assert len(a) == len(b) and all(a[k] == b[k] for k in set(a.d) | set(b.d))


# ---- mm_trickle_down (search_core.cuh) vs mm_trickle_down_warp (search_warp.cuh) --------------------------------
def trickle_seq(h, n, i, MAX):
    better = (lambda a, b: a > b) if MAX else (lambda a, b: a < b)
    e = h[i]
    while True:
        c1, g1 = 2 * i + 1, 4 * i + 3
        if c1 >= n:
            break
        best, bk, be = None, e, e
        for idx in (c1, c1 + 1, g1, g1 + 1, g1 + 2, g1 + 3):
            if idx < n and better(h[idx], bk):
                best, bk, be = idx, h[idx], h[idx]
        if best is None:
            break
        was_child = best <= c1 + 1
        h[i] = be
        i = best
        if was_child:
            break
        p = (i - 1) >> 1
        if better(h[p], e):
            h[p], e = e, h[p]
    h[i] = e


def trickle_warp(h, n, i, MAX):
    better = (lambda a, b: a > b) if MAX else (lambda a, b: a < b)
    e = h[i]

    def level(x):
        nonlocal i, e
        c1, g1 = 2 * i + 1, 4 * i + 3
        best, bk, be = None, e, e
        for c, idx in enumerate((c1, c1 + 1, g1, g1 + 1, g1 + 2, g1 + 3)):
            if idx < n and better(x[c], bk):
                best, bk, be = idx, x[c], x[c]
        if best is None:
            return None
        was_child = best <= c1 + 1
        h[i] = be
        i = best
        if was_child:
            return None
        b = best - g1
        p = (i - 1) >> 1
        pe = x[0] if b < 2 else x[1]
        if better(pe, e):
            h[p] = e
            e = pe
        return b

    while True:
        if 2 * i + 1 >= n:
            break
        v = []
        for lane in range(32):
            depth = 1 if lane < 2 else (2 if lane < 6 else (3 if lane < 14 else 4))
            off = lane if lane < 2 else (lane - 2 if lane < 6 else (lane - 6 if lane < 14 else lane - 14))
            idx = ((i + 1) << depth) - 1 + off
            v.append(h[idx] if lane < 30 and idx < n else e)
        b = level([v[c] for c in range(6)])
        if b is None:
            break
        if 2 * i + 1 >= n:
            break
        b = level([v[6 + 2 * b + c] if c < 2 else v[14 + 4 * b + (c - 2)] for c in range(6)])
        if b is None:
            break
    h[i] = e


def test_warp_trickle_down_matches_sequential():
    rng = random.Random(11)
    for trial in range(400):
        n = rng.choice([1, 2, 3, 4, 6, 7, 8, 15, 16, 31, 33, 100, 1000, 5000])
        span = rng.choice([2, 5, 1000])
        base = [(rng.randrange(span), k) for k in range(n)]
        # make roughly heap-like arrays half of the time so that long descents occur
        if rng.random() < 0.5:
            base.sort(key=lambda t: -t[0])
        for MAX in (True, False):
            for i in (0, 1, 2):
                if i >= n:
                    continue
                a, b = list(base), list(base)
                trickle_seq(a, n, i, MAX)
                trickle_warp(b, n, i, MAX)
                assert a == b, (trial, n, i, MAX)
