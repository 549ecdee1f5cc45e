"""Partition scan positions over GPUs (stripes) and mini-batches.

Host-side NumPy, bit-exact with the reference's tike.cluster
(src/tike/cluster.py:176-637): the same NumPy primitives are used for every
decision (argsort / argpartition / argmax tie-breaking, float32 means) so
that orders and batch assignments are identical index for index.  The
``compact`` method consumes NumPy's legacy global generator exactly like the
reference (cluster.py:518-532).
"""
from __future__ import annotations

import logging
import typing

import numpy as np

logger = logging.getLogger(__name__)

_UNASSIGNED = 0xFFFF


def _trivial(population, num_cluster):
    return np.array_split(np.arange(population.shape[0]), num_cluster)


def _check_count(num_cluster):
    if not 0 < num_cluster < 0xFFFF:
        raise ValueError(
            f"The number of clusters must be 0 < {num_cluster} < 65536.")


def stripes_equal_count(population, num_cluster: int, dim: int = 0):
    """Equal-count stripes along ``dim`` (cluster.py:265-299)."""
    population = np.asarray(population)
    if num_cluster == 1 or num_cluster >= len(population):
        return _trivial(population, num_cluster)
    return np.array_split(np.argsort(population[:, dim]), num_cluster)


_NATIVE_MIN_POINTS = 1024  # below this the NumPy loop is fast enough


def _grow_native(population, labels, num_cluster, start):
    """The same loop in C (libtikeb200 tb_cluster_grow), bit-exact for
    float32 (N, 2) populations; returns False when it does not apply."""
    if (population.dtype != np.float32 or population.ndim != 2
            or population.shape[1] != 2 or len(population) < _NATIVE_MIN_POINTS):
        return False
    try:
        from . import _lib
        handle = _lib.lib()
    except Exception:  # library not built: NumPy loop
        return False
    pop = np.ascontiguousarray(population)
    rc = handle.tb_cluster_grow(pop.ctypes.data, len(pop), 2, labels.ctypes.data,
                                int(num_cluster), int(start))
    if rc != 0:
        raise RuntimeError('tb_cluster_grow failed')
    return True


def _grow_heterogeneous(population, labels, num_cluster, start):
    """Round-robin: give each cluster the unlabelled point farthest from its
    current centroid (cluster.py:360-376, 447-461)."""
    if _grow_native(population, labels, num_cluster, start):
        return [np.flatnonzero(labels == c) for c in range(num_cluster)]
    for step in range(start):
        c = step % num_cluster
        free = labels == _UNASSIGNED
        centre = np.mean(population[labels == c], axis=0, keepdims=True)
        far = np.argmax(np.linalg.norm(population[free] - centre, axis=1), axis=0)
        pick = np.argmax(np.cumsum(free) == (far + 1))
        labels[pick] = c
    return [np.flatnonzero(labels == c) for c in range(num_cluster)]


def wobbly_center(population, num_cluster: int):
    """Maximally heterogeneous clusters (cluster.py:302-377)."""
    population = np.asarray(population)
    _check_count(num_cluster)
    if num_cluster == 1 or num_cluster >= len(population):
        return _trivial(population, num_cluster)
    spread = np.linalg.norm(
        population - np.mean(population, axis=0, keepdims=True), axis=1)
    seeds = np.argpartition(spread, num_cluster, axis=0)[:num_cluster]
    labels = np.full(len(population), _UNASSIGNED, dtype='uint16')
    labels[seeds] = range(num_cluster)
    return _grow_heterogeneous(population, labels, num_cluster,
                               len(population) - len(seeds))


def wobbly_center_random_bootstrap(population, num_cluster: int,
                                   boot_fraction: float = 0.95):
    """Random bootstrap followed by wobbly-center growth (cluster.py:380-462).
    Draws from NumPy's legacy global generator like the reference."""
    population = np.asarray(population)
    _check_count(num_cluster)
    if num_cluster == 1 or num_cluster >= len(population):
        return _trivial(population, num_cluster)
    num_boot = int(len(population) * boot_fraction)
    num_boot -= num_boot % num_cluster
    seed = np.random.choice(len(population), size=num_boot, replace=False)
    labels = np.full(len(population), _UNASSIGNED, dtype='uint16')
    for c in range(num_cluster):
        labels[seed[c::num_cluster]] = c
    return _grow_heterogeneous(population, labels, num_cluster,
                               len(population) - num_boot)


def compact(population, num_cluster: int, max_iter: int = 500):
    """Equal-size k-means-like clusters (cluster.py:465-637)."""
    population = np.asarray(population)
    _check_count(num_cluster)
    if num_cluster == 1 or num_cluster >= len(population):
        return _trivial(population, num_cluster)
    npts = len(population)
    everyone = np.arange(npts)
    capacity = np.full(num_cluster, npts // num_cluster)
    capacity[:npts % num_cluster] += 1
    filled = np.zeros(num_cluster, dtype='int')

    # k-means++ seeding with the legacy global generator
    seeds = np.zeros(num_cluster, dtype='int')
    seeds[0] = np.random.choice(everyone, size=1, p=None)[0]
    d2 = np.inf
    for c in range(1, num_cluster):
        d2 = np.minimum(
            d2, np.linalg.norm(population - population[seeds[c - 1]], axis=1)**2)
        seeds[c] = np.random.choice(everyone, size=1, p=d2 / d2.sum())[0]
    centroids = population[seeds]

    labels = np.full(npts, _UNASSIGNED, dtype='uint16')
    dist = np.empty((npts, num_cluster))
    open_clusters = list(range(num_cluster))
    # the reference keeps a Python list of waiting points and removes from it
    # one by one (O(n) each); the list is always in ascending order, so a mask
    # plus flatnonzero gives the same sequence
    waiting = np.ones(npts, dtype=bool)
    for c in open_clusters:
        dist[:, c] = np.linalg.norm(centroids[c] - population, axis=1)
        labels[seeds[c]] = c
        waiting[seeds[c]] = False
        filled[c] += 1
    for c in range(num_cluster):
        if filled[c] >= capacity[c]:
            open_clusters.remove(c)
    while open_clusters:
        oc = np.array(open_clusters)
        nearest = oc[np.argmin(dist[:, open_clusters], axis=1)]
        farthest = oc[np.argmax(dist[:, open_clusters], axis=1)]
        queue = np.flatnonzero(waiting)
        urgency = (dist[everyone, nearest] - dist[everyone, farthest])[queue]
        for p in queue[np.argsort(urgency)]:
            labels[p] = nearest[p]
            waiting[p] = False
            filled[nearest[p]] += 1
            if filled[nearest[p]] >= capacity[nearest[p]]:
                open_clusters.remove(nearest[p])
                break  # restart with one cluster fewer

    # pairwise swaps that improve the (heuristic) happiness
    for _ in range(max_iter):
        swapped = False
        for c in range(num_cluster):
            dist[:, c] = np.linalg.norm(centroids[c] - population, axis=1)
        wanted = np.argmin(dist, axis=1)
        happiness = dist[everyone, wanted] - dist[everyone, labels]
        native = _native_compact_sweep(dist, labels, wanted, happiness)
        if native is not None:
            swapped = native
        for p in (() if native is not None else np.argsort(happiness)):
            if happiness[p] < 0:
                gain = (dist[p, labels[p]] + dist[everyone, labels] -
                        dist[p, labels] - dist[everyone, labels[p]])
                good = np.flatnonzero(
                    np.logical_and(gain > 0, labels != labels[p]))
                if good.size > 0:
                    swapped = True
                    o = good[np.argmax(gain[good])]
                    labels[o], labels[p] = labels[p], labels[o]
                    happiness[o] = dist[o, wanted[o]] - dist[o, labels[o]]
                    happiness[p] = dist[p, wanted[p]] - dist[p, labels[p]]
        if not swapped:
            break
        for c in range(num_cluster):
            centroids[c] = np.mean(population[labels == c], axis=0)

    indices = [np.flatnonzero(labels == c) for c in range(num_cluster)]
    indices.sort(key=len, reverse=True)
    return indices


def _native_compact_sweep(dist, labels, wanted, happiness):
    """One swap sweep in libtikeb200 (host code, bit-exact with the NumPy loop
    in compact()); None when the library is not built or the problem is small."""
    if len(labels) < 1024:
        return None
    try:
        from ._lib import lib
        h = lib()
    except Exception:  # noqa: BLE001 - clustering also works without the library
        return None
    import ctypes as C
    order = np.ascontiguousarray(np.argsort(happiness), dtype=np.int64)
    wanted64 = np.ascontiguousarray(wanted, dtype=np.int64)
    assert dist.flags.c_contiguous and dist.dtype == np.float64
    assert labels.flags.c_contiguous and labels.dtype == np.uint16
    assert happiness.flags.c_contiguous and happiness.dtype == np.float64
    rc = h.tb_cluster_compact_sweep(
        dist.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p),
        wanted64.ctypes.data_as(C.c_void_p), happiness.ctypes.data_as(C.c_void_p),
        order.ctypes.data_as(C.c_void_p), len(labels), dist.shape[1])
    if rc < 0:
        raise ValueError('tb_cluster_compact_sweep: bad arguments')
    return bool(rc)


_METHODS = {
    'wobbly_center': wobbly_center,
    'wobbly_center_random_bootstrap': wobbly_center_random_bootstrap,
    'compact': compact,
}


def by_scan_stripes_contiguous(scan, num_workers: int, batch_method: str,
                               num_batch: int):
    """Stripes per worker, batches per stripe, contiguous re-indexing
    (cluster.py:176-262).

    Returns (order, batches, stripe_start): ``order[g]`` are the indices of
    the original arrays owned by worker g, already permuted so that every
    batch is a contiguous range; ``batches[g][n]`` is that range;
    ``stripe_start[g]`` is floor(min row coordinate) of the stripe.
    """
    if batch_method not in _METHODS:
        raise ValueError(f'unknown batch_method {batch_method!r}; choose from '
                         f'{sorted(_METHODS)}')
    scan = np.asarray(scan)
    owner = stripes_equal_count(scan, num_workers, dim=0)
    order: typing.List[np.ndarray] = []
    batches: typing.List[typing.List[np.ndarray]] = []
    stripe_start: typing.List[int] = []
    for mine in owner:
        o, b, s0 = stripe_batches(scan, mine, batch_method, num_batch)
        order.append(o)
        batches.append(b)
        stripe_start.append(s0)
    return order, batches, stripe_start


BAND_ROWS = 16  # same band height as kernels.PRECOND_BAND / kBand in csrc/precond.cu


def band_sort_batches(scan, order, batches):
    """Re-order the positions INSIDE every batch by (floor(row) // BAND_ROWS,
    floor(column)); batch membership, batch ranges and the batch sequence stay
    exactly what by_scan_stripes_contiguous returned.  Neighbouring positions
    are then visited back to back, which keeps their object windows and
    gradient reductions in L2 (about 2 % of the fused kernel).  Not
    in the reference: there the order inside a batch is whatever the
    clustering left, and no result depends on it beyond float summation order.

    ``order`` / ``batches`` are the first two items of the split; returns the
    new ``order`` (a list of index arrays, one per worker)."""
    scan = np.asarray(scan)
    corner = np.floor(scan).astype(np.int64)
    key = (corner[:, 0] // BAND_ROWS) * (1 << 21) + np.clip(
        corner[:, 1] + (1 << 20), 0, (1 << 21) - 1)
    out = []
    for worker_order, worker_batches in zip(order, batches):
        new = np.array(worker_order, copy=True)
        for batch in worker_batches:
            if len(batch):
                idx = new[batch]
                new[batch] = idx[np.argsort(key[idx], kind='stable')]
        out.append(new)
    return out


def boundary_mask(rows, shared_rows, probe_width: int, margin: int = 1):
    """True for the positions (given by their scan row coordinates) whose
    footprint -- rows floor(y) ... floor(y) + probe_width -- comes within
    ``margin`` rows of one of the ``shared_rows`` ranges [lo, hi) that another
    worker touches as well (communicators.RowPlan.shared_rows)."""
    top = np.floor(np.asarray(rows)).astype(np.int64)
    out = np.zeros(top.shape, dtype=bool)
    for lo, hi in shared_rows:
        out |= (top + probe_width + 1 + margin > lo) & (top - margin < hi)
    return out


def boundary_first_batches(scan, order, batches, probe_width: int, height: int):
    """Inside every batch, visit the positions next to another worker's stripe
    first (each part keeps its order, e.g. the band order of
    band_sort_batches).  The object rows two workers share are complete once
    those positions are done, so their exchange overlaps the rest of the batch
    (solvers/_common.ObjectReducer).  Batch membership and ranges are
    untouched.  Returns the new ``order``."""
    from .communicators.comm import RowPlan
    scan = np.asarray(scan)
    minmax = [(float(scan[o, 0].min()), float(scan[o, 0].max())) if len(o) else None
              for o in order]
    plan = RowPlan.from_scan_rows(minmax, probe_width, height)
    out = []
    for w, (worker_order, worker_batches) in enumerate(zip(order, batches)):
        new = np.array(worker_order, copy=True)
        shared = plan.shared_rows(w)
        for batch in worker_batches:
            if len(batch) and shared:
                idx = new[batch]
                inner = ~boundary_mask(scan[idx, 0], shared, probe_width)
                new[batch] = idx[np.argsort(inner, kind='stable')]
        out.append(new)
    return out


def stripe_batches(scan, mine, batch_method: str, num_batch: int):
    """Batches of ONE stripe (the body of the loop in
    by_scan_stripes_contiguous): (order, batches, stripe_start) of the worker
    owning the positions ``mine``.  ``wobbly_center`` is deterministic, so in a
    multi-process run every rank can compute just its own stripe."""
    local = np.asarray(scan[mine], dtype=scan.dtype)
    start = int(np.floor(np.min(local[:, 0])))
    groups = _METHODS[batch_method](local, num_cluster=num_batch)
    order = mine[np.concatenate(groups)]
    breaks = np.cumsum([len(g) for g in groups])[:-1]
    return order, np.array_split(np.arange(len(mine)), breaks), start


def by_scan_stripes(scan, n: int, fly: int = 1, axis: int = 0):
    """``n`` boolean masks splitting the field of view into equal-width stripes
    along ``axis`` (cluster.py:123-173)."""
    if scan.ndim != 2:
        raise ValueError('scan must have shape (nscan, 2)')
    nscan = scan.shape[0]
    if nscan % fly != 0:
        raise ValueError('The number of scan positions must be divisible by fly')
    coord = scan.reshape(nscan // fly, fly, scan.shape[-1])[:, 0, axis]
    edges = np.linspace(coord.min(), coord.max(), n + 1, endpoint=True)
    edges[0] -= 1  # widen the outer stripes so every point is claimed
    edges[-1] += 1
    return [np.logical_and(edges[i] < coord, coord <= edges[i + 1]).repeat(fly)
            for i in range(n)]


def cluster_wobbly_center(*args, **kwargs):
    """Deprecated alias of wobbly_center (cluster.py:663-670)."""
    import warnings
    warnings.warn('cluster_wobbly_center is deprecated. Use wobbly_center instead.',
                  DeprecationWarning)
    return wobbly_center(*args, **kwargs)


def cluster_compact(*args, **kwargs):
    """Deprecated alias of compact (cluster.py:673-680)."""
    import warnings
    warnings.warn('cluster_compact is deprecated. Use compact instead.',
                  DeprecationWarning)
    return compact(*args, **kwargs)
