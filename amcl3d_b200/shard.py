"""Host-side sharding logic of the multi-GPU path, as plain numpy (one process per GPU, particles block-partitioned,
grid replicated).  The CUDA library implements exactly this algebra (csrc/filter.cu: update_fast_stage1/2,
resample_gather); this module is its executable specification, used by bench.py for partitioning and by the
world_size-2 gloo tests on CPU.
"""
import numpy as np


def partition(n_total, rank, world):
    """Block partition by particle index (keeps the resample order): returns (first, count)."""
    per = (n_total + world - 1) // world
    first = min(n_total, rank * per)
    return first, min(n_total, first + per) - first


def update_partials(poses_xyza, wp, wr, inside):
    """The ten fp64 partial sums one rank contributes to an update (only in-map particles count):
    [sum wp, sum wr, sum wp*x, wp*y, wp*z, wp*a, sum wr*x, wr*y, wr*z, wr*a]."""
    m = np.asarray(inside, bool)
    p = np.asarray(poses_xyza, np.float64)[m]
    a = np.asarray(wp, np.float64)[m]
    b = np.asarray(wr, np.float64)[m]
    out = np.zeros(10)
    out[0], out[1] = a.sum(), b.sum()
    out[2:6] = (a[:, None] * p).sum(0)
    out[6:10] = (b[:, None] * p).sum(0)
    return out


def finish_update(partials, wp, wr, inside, alpha):
    """After the all-reduce of the partials: normalised wp / wr / w for the local particles and the global mean.
    Folds the reference's three dependent sums (ParticleFilter.cpp:151-152,179,190-193) into one reduction:
    the sum over in-map particles of wp/sum(wp) is 1, so wt = alpha*[sum wp > 0] + (1-alpha)*[sum wr > 0]."""
    A, B = partials[0], partials[1]
    wtp, wtr = np.float32(A), np.float32(B)
    wt_d = (alpha if A > 0 else 0.0) + ((1.0 - alpha) if B > 0 else 0.0)
    wt = np.float32(wt_d)
    wp = np.asarray(wp, np.float32)
    wr = np.asarray(wr, np.float32)
    wpn = (wp / wtp).astype(np.float32) if wtp > 0 else np.zeros_like(wp)
    wrn = (wr / wtr).astype(np.float32) if wtr > 0 else np.zeros_like(wr)
    w = (wpn.astype(np.float64) * alpha + wrn.astype(np.float64) * (1.0 - alpha)).astype(np.float32)
    w = np.where(np.asarray(inside, bool), w, np.float32(0))
    w = (w / wt).astype(np.float32) if wt > 0 else np.zeros_like(w)
    mean = np.zeros(4)
    if wt_d > 0:
        if A > 0:
            mean += alpha * partials[2:6] / A
        if B > 0:
            mean += (1.0 - alpha) * partials[6:10] / B
        mean /= wt_d
    return wpn, wrn, w, mean.astype(np.float32)


def resample_indices(weights_all, u01, first, count):
    """Scan-mode global resample for the output slots [first, first+count): fp64 inclusive prefix over ALL ranks'
    weights, u_m = r + factor*m in float (ParticleFilter.cpp:201-209), first index whose prefix reaches u_m,
    clamped to n-1."""
    w = np.asarray(weights_all, np.float32)
    n = len(w)
    factor = np.float32(1.0) / np.float32(n)
    r = np.float32(factor * np.float32(u01))
    m = np.arange(first, first + count, dtype=np.uint32).astype(np.float32)
    u = (r + (factor * m).astype(np.float32)).astype(np.float32)
    cdf = np.cumsum(w.astype(np.float64))
    return np.minimum(np.searchsorted(cdf, u.astype(np.float64), side="left"), n - 1).astype(np.uint32)


def slab_boundaries(layer_points, pad_z, tz_total, tiles_xy, n_ranks, row_tiles=1):
    """z-slab boundaries of the sharded computeGrid (csrc/distance_field.cu), in z-tiles: slab r = [b[r], b[r+1]).
    `layer_points[bz]` = map points bucketed into block layer bz (tile tz lives in block layer tz + pad_z).  A tile
    layer costs 20 point-visits per tile plus one per point in its own and the two adjacent block layers; rows of
    `row_tiles` layers (a brick row when the grid is bricked) are handed out so that every rank's cost reaches its
    share of the total.  Identical on every rank because every rank buckets the same points."""
    layer_points = np.asarray(layer_points, np.float64)
    n_rows = (tz_total + row_tiles - 1) // row_tiles
    cost = np.zeros(n_rows)
    for tz in range(tz_total):
        pts = 0.0
        for dz in (-1, 0, 1):
            bz = tz + pad_z + dz
            if 0 <= bz < len(layer_points):
                pts += layer_points[bz]
        cost[tz // row_tiles] += 20.0 * tiles_xy + pts
    total = cost.sum()
    b = [0] + [tz_total] * n_ranks
    acc, r = 0.0, 1
    for row in range(n_rows):
        if r >= n_ranks:
            break
        acc += cost[row]
        while r < n_ranks and acc >= total * r / n_ranks:
            b[r] = min(tz_total, (row + 1) * row_tiles)
            r += 1
    return b


def deal_chunks(sorted_order, world, chunk):
    """The weighting schedule of a sharded particle set (csrc/filter.cu: refresh_global_schedule, deal_chunks_kernel):
    the pose-sorted permutation of ALL particles is cut into chunks of `chunk` particles, chunk c goes to region c % world
    at position c // world, and rank r then weighs the r-th slice of ceil(n / world) entries of the result.  chunk = 0 (or a
    set no longer than chunk * world) keeps contiguous slices.  Returns the dealt permutation."""
    order = np.asarray(sorted_order)
    n = len(order)
    if chunk <= 0 or n <= chunk * world:
        return order.copy()
    n_chunks = (n + chunk - 1) // chunk
    start = np.zeros(world, np.int64)
    at = 0
    for r in range(world):
        start[r] = at
        mine = (n_chunks - 1 - r) // world + 1 if n_chunks > r else 0
        particles = mine * chunk
        if mine and (n_chunks - 1) % world == r:
            particles -= n_chunks * chunk - n          # the very last chunk of the permutation is the only partial one
        at += particles
    i = np.arange(n, dtype=np.int64)
    c, o = i // chunk, i % chunk
    dealt = np.empty_like(order)
    dealt[start[c % world] + (c // world) * chunk + o] = order
    return dealt
