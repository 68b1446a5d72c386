"""CPU: exhaustive model check of the peer-memory exchange protocol of the sharded update (csrc/comm.cu PeerBox,
csrc/filter.cu update_fast_stage1/2_kernel).

Protocol: in step k every rank r stores its partial sums into slot [k & 1][r] of EVERY rank's mailbox and then raises
the flag there to k (one store per destination, in any order, at any time relative to the other ranks); stage 2 of
rank r waits until all flags [k & 1][*] of ITS mailbox equal k, then adds the slots.  Step k + 1 of a rank starts
only after its stage 2 of step k (stream order).  Claim: two slots per source (step parity) suffice -- when a rank
reads step k no slot can already hold step k + 2 -- and no interleaving deadlocks.

The model explores every interleaving of the per-destination stores and the reads for 2 and 3 ranks."""
from collections import deque

import pytest


def explore(n_ranks, n_steps):
    # state: per rank (step, tuple of destinations still to be written in this step or None when waiting to read),
    #        mailbox[dst][parity][src] = last step whose flag was stored there (0 = never)
    all_dst = tuple(range(n_ranks))
    start_ranks = tuple((1, all_dst) for _ in range(n_ranks))
    start_box = tuple(tuple(tuple(0 for _ in range(n_ranks)) for _ in range(2)) for _ in range(n_ranks))
    seen = {(start_ranks, start_box)}
    todo = deque(seen)
    finished = False
    while todo:
        ranks, box = todo.popleft()
        if all(step > n_steps for step, _ in ranks):
            finished = True
            continue
        moved = False
        for r, (step, pending) in enumerate(ranks):
            if step > n_steps:
                continue
            if pending:                                   # stage 1, last CTA: one store per destination, any order
                for dst in pending:
                    par = step & 1
                    old = box[dst][par][r]
                    # the slot being overwritten must have been consumed: the destination is past step - 2
                    assert old in (0, step - 2), "slot overwritten out of order"
                    assert old == 0 or ranks[dst][0] > step - 2, "step %d overwrites an unread step %d" % (step, old)
                    nb = [list(map(list, b)) for b in box]
                    nb[dst][par][r] = step
                    nbox = tuple(tuple(tuple(p) for p in b) for b in nb)
                    rest = tuple(d for d in pending if d != dst)
                    nranks = ranks[:r] + ((step, rest),) + ranks[r + 1:]
                    moved = True
                    if (nranks, nbox) not in seen:
                        seen.add((nranks, nbox))
                        todo.append((nranks, nbox))
            else:                                         # stage 2: spin until every flag of the own box shows `step`
                flags = box[r][step & 1]
                assert all(f in (step - 2, step) or (f == 0 and step <= 2) for f in flags), "a flag ran ahead: %r" % (flags,)
                if all(f == step for f in flags):
                    nranks = ranks[:r] + ((step + 1, all_dst),) + ranks[r + 1:]
                    moved = True
                    if (nranks, box) not in seen:
                        seen.add((nranks, box))
                        todo.append((nranks, box))
        assert moved, "deadlock: %r" % (ranks,)
    return finished, len(seen)


@pytest.mark.parametrize("n_ranks,n_steps", [(2, 6), (3, 4)])
def test_two_parity_slots_suffice_and_nothing_deadlocks(n_ranks, n_steps):
    finished, states = explore(n_ranks, n_steps)
    assert finished and states > 100


def test_a_single_slot_would_not_suffice():
    """Sanity of the model itself: with ONE slot per source (no parity) a fast rank overwrites a value its peer has
    not read yet -- the checker must notice."""
    n = 2
    all_dst = (0, 1)
    ranks = ((1, all_dst), (1, all_dst))
    box = [[0, 0], [0, 0]]          # box[dst][src], single slot
    # rank 0 publishes step 1 everywhere, rank 1 publishes step 1 everywhere, rank 0 reads (both flags 1) and moves to
    # step 2, publishes step 2 into rank 1's box BEFORE rank 1 has read step 1
    for r in (0, 1):
        for d in all_dst:
            box[d][r] = 1
    assert box[0] == [1, 1]
    box[1][0] = 2                    # rank 0, step 2 -> rank 1's only slot for source 0
    assert box[1] != [1, 1], "rank 1 can no longer read step 1: a single slot loses data"
