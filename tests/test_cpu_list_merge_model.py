"""CPU model of the candidate-list merge of the search kernels (csrc/select.cuh, `list_merge`).

A row's shared-memory record is `cap` sorted entries (padded with EMPTY = key +inf / column -1
while the list is not full) followed by `cnt` unsorted appended entries; entries are packed
(order-preserving key bits << 32 | column), unique.  The kernel merges BY RANK: every entry
computes the position it has in the sorted union and is scattered there; positions below `cap`
must be written exactly once: a list entry moves up by the number of buffered entries below it, a
buffered entry lands at (list entries below it, counted with one ballot) + (buffered entries
below it).  This restates that arithmetic in Python, lane by lane, and checks it against
`sorted()` over list lengths, buffer sizes, fill levels and tied keys -- the property the GPU
tests rely on through the kNN parity checks."""
import random

import pytest

EMPTY = (0xFF800000 << 32) | 0xFFFFFFFF


def rank_merge(e, cap, cnt):
    """select.cuh list_merge_rank, lane by lane: returns the new list; asserts that the scatter is
    a permutation.  NL / NB = list / buffer entries per lane, as the kernel dispatches them."""
    nl = {1: 1, 2: 2, 3: 4, 4: 4}[(cap + 31) // 32]
    nb = 1 if cnt <= 32 else 2
    buf = e[cap:cap + cnt]
    lanes = range(32)
    vl = {(ln, t): ln + 32 * t < cap for ln in lanes for t in range(nl)}
    xl = {(ln, t): e[ln + 32 * t] if vl[ln, t] else EMPTY for ln in lanes for t in range(nl)}
    pl = {(ln, t): ln + 32 * t if vl[ln, t] else 1 << 30 for ln in lanes for t in range(nl)}
    xb = {(ln, u): buf[ln + 32 * u] if ln + 32 * u < cnt else EMPTY for ln in lanes for u in range(nb)}
    pb = {(ln, u): 0 if ln + 32 * u < cnt else 1 << 30 for ln in lanes for u in range(nb)}
    for j in range(cnt):                                  # entry j is broadcast to every lane
        y = buf[j]
        below = 0                                         # list entries below y: one ballot per t
        for t in range(nl):
            up = {ln: y < xl[ln, t] for ln in lanes}      # y lands below my list entry
            for ln in lanes:
                pl[ln, t] += 1 if up[ln] else 0
            below += sum(1 for ln in lanes if vl[ln, t] and not up[ln])
        for ln in lanes:
            for u in range(nb):
                pb[ln, u] += 1 if y < xb[ln, u] else 0
                if ln == (j & 31) and u == (j >> 5):      # the lane that owns entry j
                    pb[ln, u] += below
    written = {}
    for pos, value in [(pl[k], xl[k]) for k in pl] + [(pb[k], xb[k]) for k in pb]:
        if pos < cap:
            assert pos not in written, "two entries scattered to one position"
            written[pos] = value
    assert sorted(written) == list(range(cap)), "a list position was not written"
    return [written[p] for p in range(cap)]


@pytest.mark.parametrize("cap", [8, 16, 24, 32, 56, 64, 112, 128])
@pytest.mark.parametrize("slots", [8, 12, 20, 28, 44, 64])
def test_rank_merge_equals_sorted_union(cap, slots):
    rnd = random.Random(cap * 131 + slots)
    for _ in range(40):
        real = rnd.randint(0, cap)                        # entries the list holds so far
        key_range = rnd.choice([3, 1000, 1 << 30])        # small ranges force tied keys
        cols = rnd.sample(range(1 << 20), real + slots)
        ents = [(rnd.randint(0, key_range) << 32) | c for c in cols]
        lst = sorted(ents[:real]) + [EMPTY] * (cap - real)
        cnt = rnd.randint(0, slots)
        record = lst + ents[real:real + cnt] + [0xDEAD] * (slots - cnt)   # stale slots past cnt
        got = rank_merge(record, cap, cnt)
        assert got == sorted(lst + ents[real:real + cnt])[:cap]
        assert got[cap - 1] >= got[0]


def test_rank_merge_keeps_the_lower_column_among_tied_keys():
    """sklearn's heap rejects val >= heap_max (utils/_heap.pyx), i.e. among equal keys the
    earlier (lower) column stays: packed entries order ties by column."""
    cap = 4
    lst = [(5 << 32) | 10, (7 << 32) | 3, (7 << 32) | 9, (9 << 32) | 1]
    buf = [(7 << 32) | 5, (9 << 32) | 0]
    got = rank_merge(lst + buf, cap, 2)
    assert got == [(5 << 32) | 10, (7 << 32) | 3, (7 << 32) | 5, (7 << 32) | 9]
