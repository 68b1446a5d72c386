"""CPU: the cost-balanced z-slab partition of the sharded computeGrid (amcl3d_b200/shard.py::slab_boundaries, the
executable specification of the boundaries csrc/distance_field.cu computes)."""
import numpy as np
import pytest

from amcl3d_b200 import shard


def slab_costs(layer_points, pad_z, tz_total, tiles_xy, b):
    lp = np.asarray(layer_points, np.float64)
    out = []
    for r in range(len(b) - 1):
        c = 0.0
        for tz in range(b[r], b[r + 1]):
            c += 20.0 * tiles_xy
            for dz in (-1, 0, 1):
                bz = tz + pad_z + dz
                if 0 <= bz < len(lp):
                    c += lp[bz]
        out.append(c)
    return np.array(out)


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("row_tiles", [1, 4])
def test_boundaries_cover_the_grid_and_balance_a_warehouse(n_ranks, row_tiles):
    # map L: 50 tile layers of 0.4 m; floor, shelving up to 8 m (20 layers), ceiling; 250 x 250 tiles per layer
    tz_total, pad_z, tiles_xy = 50, 1, 62500
    lp = np.zeros(tz_total + 2 * pad_z)
    lp[pad_z:pad_z + 21] = 1.3e6
    lp[pad_z] += 4e6
    lp[pad_z + 49] = 4e6
    b = shard.slab_boundaries(lp, pad_z, tz_total, tiles_xy, n_ranks, row_tiles)
    assert b[0] == 0 and b[-1] == tz_total and len(b) == n_ranks + 1
    assert all(b[i] <= b[i + 1] for i in range(n_ranks))
    assert all(x % row_tiles == 0 or x == tz_total for x in b)
    costs = slab_costs(lp, pad_z, tz_total, tiles_xy, b)
    equal = slab_costs(lp, pad_z, tz_total, tiles_xy,
                       [min(tz_total, r * -(-tz_total // n_ranks)) for r in range(n_ranks)] + [tz_total])
    # never worse than equal-height slabs, and within one row of the ideal share
    row_cost_max = max(slab_costs(lp, pad_z, tz_total, tiles_xy, [k, min(tz_total, k + row_tiles)])[0]
                       for k in range(0, tz_total, row_tiles))
    assert costs.max() <= equal.max() + 1e-6
    assert costs.max() <= costs.sum() / n_ranks + row_cost_max + 1e-6


def test_more_ranks_than_rows_leaves_trailing_slabs_empty():
    b = shard.slab_boundaries(np.ones(6) * 100, 1, 4, 10, 8, row_tiles=4)
    assert b[0] == 0 and b[-1] == 4
    assert sum(1 for i in range(8) if b[i + 1] > b[i]) == 1


def test_empty_map_splits_by_height():
    b = shard.slab_boundaries(np.zeros(12), 1, 10, 100, 5)
    assert b == [0, 2, 4, 6, 8, 10]
