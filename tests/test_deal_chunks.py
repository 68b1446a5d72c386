"""CPU: the chunk dealing of the sharded weighting schedule (amcl3d_b200/shard.py: deal_chunks, the specification of
deal_chunks_kernel in csrc/filter.cu) is a permutation for every size / rank count / chunk, keeps chunks intact and gives
every rank every world-th chunk."""
import numpy as np
import pytest

from amcl3d_b200 import shard


@pytest.mark.parametrize("n,world,chunk", [(8229, 2, 2048), (8229, 2, 16384), (1048576, 8, 16384), (1048577, 8, 16384),
                                           (100000, 3, 4096), (65536 * 3 + 5, 3, 65536), (12345, 5, 1000),
                                           (4096, 4, 1024), (4097, 4, 1024), (10, 8, 1)])
def test_dealing_is_a_permutation_that_keeps_chunks(n, world, chunk):
    order = np.random.default_rng(n + world).permutation(n).astype(np.uint32)
    dealt = shard.deal_chunks(order, world, chunk)
    assert sorted(dealt.tolist()) == list(range(n))
    if n <= chunk * world:
        assert np.array_equal(dealt, order)
        return
    # every chunk of the sorted order appears contiguously in the dealt order
    pos = np.empty(n, np.int64)
    pos[dealt] = np.arange(n)
    for c in range((n + chunk - 1) // chunk):
        p = pos[order[c * chunk:(c + 1) * chunk]]
        assert np.array_equal(p, np.arange(p[0], p[0] + len(p)))
    # region r (what rank r weighs, up to the ragged ends) holds the chunks c = r, r + world, ... in that order
    n_chunks = (n + chunk - 1) // chunk
    first_of = [pos[order[c * chunk]] for c in range(n_chunks)]
    for r in range(world):
        mine = [first_of[c] for c in range(r, n_chunks, world)]
        assert mine == sorted(mine)
        if r + 1 < world and n_chunks > r + 1:
            assert max(mine) < min(first_of[c] for c in range(r + 1, n_chunks, world))


def test_slices_cover_every_particle_once():
    n, world = 1048576 + 37, 8
    dealt = shard.deal_chunks(np.arange(n), world, 16384)
    seen = np.zeros(n, np.int32)
    per = (n + world - 1) // world
    for r in range(world):
        first = min(n, r * per)
        seen[dealt[first:min(n, first + per)]] += 1
    assert np.all(seen == 1)
