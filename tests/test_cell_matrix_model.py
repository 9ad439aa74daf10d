"""The arithmetic behind the sliced build's slot image (csrc/sdt_skm.cuh), stated as a model and checked
against the reference's update_kmer (newhash.c:71-96) on random instance streams.

A slot holds a 5 x 5 matrix of (left, right) cells, 4 = "no neighbour", plus an overflow count `extra`:
an instance with multiplicity `add` (skm_dedupe_kernel) goes into its cell as long as the cell can still
matter to a 6-bit link counter and into `extra` otherwise.  On the way out
    count = sum of all cells + extra (mod 2^32),  L[b] = min (63, row b),  R[b] = min (63, column b).
The model applies the instances in any order and in any interleaving of the "read the cell, then add"
steps (the kernel's lanes race on exactly that), and must always land on the reference's counters."""
import numpy as np
import pytest

LINK_SAT, CELL_STOP = 63, 0xF000


def reference(instances):
    count, L, R = 0, [0] * 4, [0] * 4
    for left, right, add in instances:
        for _ in range(add):                       # update_kmer is applied once per instance
            count = (count + 1) & 0xFFFFFFFF
            if left < 4 and L[left] < LINK_SAT:
                L[left] += 1
            if right < 4 and R[right] < LINK_SAT:
                R[right] += 1
    return count, L, R


def image(instances, rng, stale):
    """`stale`: how many other updates may slip in between a lane's read of the cell and its add."""
    cell, extra = np.zeros((5, 5), dtype=np.int64), 0
    pending = []                                   # (left, right, add, cell value seen)
    def commit(left, right, add, seen):
        nonlocal extra
        if add == 1:                               # the `one` path: count in the cell until it nears 16 bits
            if seen >= CELL_STOP:
                extra += 1
            else:
                cell[left, right] += 1
        else:                                      # a multiplicity: what can still matter to a link counter, the rest to extra
            inc = 0 if seen >= LINK_SAT else min(add, LINK_SAT)
            cell[left, right] += inc
            extra += add - inc
    for left, right, add in instances:
        pending.append((left, right, add, int(cell[left, right])))
        while len(pending) > stale or (pending and rng.random() < 0.5):
            commit(*pending.pop(rng.integers(len(pending))))
    while pending:
        commit(*pending.pop(rng.integers(len(pending))))
    assert cell.max() < 1 << 16                    # a cell is a 16-bit field of a shared-memory word
    count = (int(cell.sum()) + extra) & 0xFFFFFFFF
    L = [min(LINK_SAT, int(cell[b, :].sum())) for b in range(4)]
    R = [min(LINK_SAT, int(cell[:, b].sum())) for b in range(4)]
    return count, L, R


@pytest.mark.parametrize("seed", range(12))
def test_cell_matrix_equals_update_kmer(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    hot = rng.random() < 0.5
    inst = []
    for _ in range(n):
        left, right = (int(x) for x in rng.integers(0, 5, size=2))
        if hot:                                    # one dominant (left, right) pair, as at a highly expressed locus
            if rng.random() < 0.8:
                left, right = 1, 3
        add = 1 if rng.random() < 0.6 else int(rng.integers(2, 300))
        inst.append((left, right, add))
    want = reference(inst)
    for stale in (0, 7, 1024):                     # up to one in-flight update per thread of the CTA
        order = [inst[i] for i in rng.permutation(n)]
        assert image(order, rng, stale) == want
