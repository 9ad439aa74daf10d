"""Models of three pieces of device-side arithmetic of the sliced build, checked on the CPU:

* `mg_swz` (csrc/sdt_merge.cuh): where the 16-byte parts of a staged record sit in shared memory.  The claim
  in the kernel's comment — consecutive threads that read the same part of consecutive records cover all 32
  banks with every quarter-warp — is checked for the three record sizes, and so is that the map is a bijection.
* the greedy group plan of `skm_merge_kernel` (warp 0's scan over the next chains' record counts): groups are
  maximal prefixes that fit a chunk, every chain is in exactly one group, order is kept.
* the end-of-run bits of `skm_emit_kernel` (step 2 writes them per group of four windows, step 3 reads them a
  segment of 16 at a time and traces a run back to the end before it): the runs reported must be exactly
  the maximal runs of equal minimizer values of a read.
"""
import numpy as np
import pytest

BANKS, BANK_BYTES = 32, 4


def mg_swz(W, rec, part):
    vec = {1: 2, 2: 3, 4: 4}[W]
    if W == 1:
        return rec * vec + (part ^ ((rec >> 2) & 1))
    if W == 4:
        return rec * vec + (part ^ ((rec >> 1) & 3))
    return rec * vec + part


@pytest.mark.parametrize("W", [1, 2, 4])
def test_swizzle_is_a_bijection_and_conflict_free(W):
    vec = {1: 2, 2: 3, 4: 4}[W]
    n = 2048 if W == 1 else 1024
    seen = {mg_swz(W, r, p) for r in range(n) for p in range(vec)}
    assert seen == set(range(n * vec))                      # every 16-byte cell of the chunk is used exactly once
    for r in range(n):                                      # a record's parts stay inside the record's own bytes
        assert {mg_swz(W, r, p) // vec for p in range(vec)} == {r}
    # a 128-bit shared access is served a quarter-warp (8 threads x 16 B = 128 B = all banks) at a time: the 8
    # threads must touch 8 different 16-byte bank groups
    for first in range(0, n - 8, 8):
        for part in range(vec):
            groups = {(mg_swz(W, first + t, part) * 16 // BANK_BYTES % BANKS) // 4 for t in range(8)}
            assert len(groups) == 8, (W, first, part)


def plan_groups(counts, chunk, gmax=32):
    """What warp 0 computes: from chain c0 on, the longest run of consecutive chains (at most gmax) whose record
    counts sum to at most `chunk`; at least one chain (a chain longer than a chunk goes through in several chunks)."""
    groups, c0 = [], 0
    while c0 < len(counts):
        window = counts[c0:c0 + gmax]
        incl = np.cumsum(np.minimum(window, 0x4000000))
        k = max(1, int((incl <= chunk).sum()))
        k = min(k, len(window))
        groups.append((c0, k, int(incl[k - 1])))
        c0 += k
    return groups


@pytest.mark.parametrize("seed", range(4))
def test_greedy_groups_partition_the_chains(seed):
    rng = np.random.default_rng(seed)
    counts = rng.geometric(1 / 357.0, size=5000)            # C2: 357 records per chain on average
    counts[rng.integers(0, 5000, 40)] = rng.integers(2049, 20000, 40)     # a few chains longer than a chunk
    counts[rng.integers(0, 5000, 200)] = 0                   # and empty ones
    groups = plan_groups(counts, 2048)
    assert [g[0] for g in groups] == list(np.cumsum([0] + [g[1] for g in groups[:-1]]))      # consecutive, in order
    assert sum(g[1] for g in groups) == len(counts)
    for c0, k, total in groups:
        assert 1 <= k <= 32
        assert total == int(np.minimum(counts[c0:c0 + k], 0x4000000).sum())
        if k > 1:
            assert total <= 2048                            # several chains only if they fit one chunk together
        if c0 + k < len(counts) and k < 32 and total <= 2048:
            assert total + counts[c0 + k] > 2048            # maximal: the next chain would not have fitted
    fill = np.mean([t for _, _, t in groups if t <= 2048]) / 2048
    assert fill > 0.75                                      # (two chains per group, the old rule, filled 35 %)


def runs_by_end_bits(values):
    """Steps 2 and 3 of skm_emit_kernel for one read: per group of four windows a byte of "a run ends at window q"
    bits (bit q), a word per 16 windows; the thread of a segment reports the runs that END in it and finds the
    start of the first one behind the last end bit of the words before its own."""
    n = len(values)
    spr = (n + 15) // 16
    endb = np.zeros(spr * 4, dtype=np.uint8)
    for j0 in range(0, n, 4):
        eb = 0
        for q in range(4):
            j = j0 + q
            if j < n and (j + 1 >= n or values[j + 1] != values[j]):
                eb |= 1 << q
        endb[j0 >> 2] = eb
    words = endb.view("<u4")
    runs = []
    for seg in range(spr):
        ends, start = int(words[seg]), 0
        for w in range(seg - 1, -1, -1):
            e2 = int(words[w])
            if e2:
                p = e2.bit_length() - 1
                start = 16 * w + ((p >> 3) << 2) + (p & 7) + 1
                break
        while ends:
            p = (ends & -ends).bit_length() - 1
            ends &= ends - 1
            j = 16 * seg + ((p >> 3) << 2) + (p & 7)
            runs.append((start, j - start + 1))
            start = j + 1
    return runs


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 70, 120, 136])
def test_end_bits_give_the_maximal_runs(n):
    rng = np.random.default_rng(n)
    for trial in range(50):
        p_change = rng.choice([0.0, 0.02, 0.12, 0.5, 1.0])
        values = np.cumsum(rng.random(n) < p_change)        # runs of equal values, average length 1 / p_change
        want, s = [], 0
        for j in range(n):
            if j + 1 == n or values[j + 1] != values[j]:
                want.append((s, j - s + 1))
                s = j + 1
        assert runs_by_end_bits(values) == want
