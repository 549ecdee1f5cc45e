"""world_size-2 gloo tests of the multi-GPU host logic (runs on CPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def _collectives(rank, world):
    from tike_b200.communicators import Comm
    comm = Comm()
    assert comm.size == world and comm.rank == rank
    t = torch.full((3,), float(rank + 1))
    comm.allreduce_sum_(t)
    z = torch.full((2, 2), complex(rank + 1, -rank), dtype=torch.complex64)
    comm.allreduce_sum_(z)
    m = torch.full((2,), float(rank))
    comm.allreduce_mean_(m)
    b = torch.full((2,), float(rank + 5))
    comm.bcast_(b, src=0)
    objs = comm.allgather_object({'rank': rank})
    s = comm.reduce_cpu_sum(np.array([1.0, rank]))
    return (t.tolist(), z[0, 0].item(), m.tolist(), b.tolist(),
            [o['rank'] for o in objs], s.tolist())


def test_comm_collectives_gloo():
    res = _spawn(_collectives)
    for t, z, m, b, ranks, s in res:
        assert t == [3.0, 3.0, 3.0]
        assert z == complex(3, -1)
        assert m == [0.5, 0.5]
        assert b == [5.0, 5.0]
        assert ranks == [0, 1]
        assert s == [2.0, 1.0]


def _swap(rank, world):
    from tike_b200.communicators import Comm
    comm = Comm()
    psi = torch.full((1, 16, 4), complex(rank), dtype=torch.complex64)
    comm.swap_edges(psi, overlap=3, edges=[0, 8])
    return psi[0, :, 0].real.tolist()


def test_swap_edges_blends_like_reference():
    """pool.py:415-476: band [edge, edge+overlap) becomes
    rampd * lower + rampu * upper on BOTH neighbours."""
    lower, upper = _spawn(_swap)
    ramp = np.linspace(0, 1, 5)[1:-1]
    expect = (1 - ramp) * 0 + ramp * 1
    np.testing.assert_allclose(lower[8:11], expect, atol=1e-6)
    np.testing.assert_allclose(upper[8:11], expect, atol=1e-6)
    assert lower[:8] == [0.0] * 8 and upper[11:] == [1.0] * 5


def test_partition_is_consistent_across_ranks():
    """Every rank derives the same stripes; union covers all positions."""
    from tike_b200 import cluster
    rng = np.random.default_rng(2)
    scan = (rng.random((203, 2)) * 100).astype(np.float32)
    order, batches, start = cluster.by_scan_stripes_contiguous(scan, 2, 'wobbly_center', 3)
    assert np.array_equal(np.sort(np.concatenate(order)), np.arange(203))
    assert scan[order[0], 0].max() <= scan[order[1], 0].min()
    assert start == [int(np.floor(scan[o, 0].min())) for o in order]
    assert all(len(b) == 3 for b in batches)


def test_stitch_stripes():
    from tike_b200.communicators import stitch_stripes
    parts = [np.full((1, 20, 3), i, np.complex64) for i in range(3)]
    out = stitch_stripes(parts, stripe_start=[0, 6, 12], probe_width=4)
    col = out[0, :, 0].real
    assert list(col[:8]) == [0] * 8 and list(col[8:14]) == [1] * 6 and list(col[14:]) == [2] * 6


def _halo(rank, world):
    """Row-bounded sum: every rank contributes on the rows it touches; after
    halo_sum_ the touched rows hold the sum over all ranks, and after an
    owner-side update gather_owned_rows_ makes every replica identical."""
    from tike_b200.communicators import Comm, RowPlan
    comm = Comm()
    H, W, N = 40, 5, 6
    # three stripes whose footprints overlap their neighbours (and, for the
    # thin middle stripe, the neighbour after next)
    minmax = [(0.3, 11.9), (12.2, 14.7), (15.1, 30.5)][:world] if world == 3 else [(0.3, 17.9), (18.2, 30.5)]
    plan = RowPlan.from_scan_rows(minmax, N, H)
    rng = np.random.default_rng(7)
    full = [rng.standard_normal((2, H, W)) + 1j * rng.standard_normal((2, H, W)) for _ in range(world)]
    contrib = []
    for r in range(world):
        c = np.zeros((2, H, W), np.complex64)
        lo, hi = plan.touched[r]
        c[:, lo:hi] = full[r][:, lo:hi]
        contrib.append(c)
    total = np.sum(contrib, axis=0)
    t = torch.from_numpy(contrib[rank].copy())
    comm.halo_sum_(t, plan)
    lo, hi = plan.touched[rank]
    err_touched = float(np.abs(t.numpy()[:, lo:hi] - total[:, lo:hi]).max())
    # owner-side state, then replicas everywhere
    psi = torch.full((2, H, W), complex(-1, -1), dtype=torch.complex64)
    olo, ohi = plan.own(rank)
    psi[:, olo:ohi] = torch.from_numpy(total[:, olo:ohi].astype(np.complex64))
    # first only the halo rows (what the next epoch reads) ...
    comm.halo_refresh_(psi, plan)
    err_refresh = float(np.abs(psi.numpy()[:, lo:hi] - total[:, lo:hi]).max())
    alo, ahi = plan.active(rank)
    assert alo <= lo and ahi >= hi
    # ... then complete replicas
    comm.gather_owned_rows_(psi, plan)
    err_gather = max(err_refresh, float(np.abs(psi.numpy() - total).max()))
    mx = torch.tensor([float(rank)])
    comm.allreduce_max_(mx)
    return err_touched, err_gather, plan.bounds, float(mx)


def test_halo_sum_and_owned_row_gather_two_ranks():
    for err_touched, err_gather, bounds, mx in _spawn(_halo, world=2):
        assert err_touched < 1e-6 and err_gather < 1e-6
        assert bounds == [0, 18, 40] and mx == 1.0


def test_halo_sum_thin_middle_stripe_three_ranks():
    for err_touched, err_gather, bounds, mx in _spawn(_halo, world=3):
        assert err_touched < 1e-6 and err_gather < 1e-6
        assert bounds == [0, 12, 15, 40] and mx == 2.0


def test_row_plan_partition():
    from tike_b200.communicators import RowPlan
    plan = RowPlan.from_scan_rows([(1.5, 9.0), None, (9.5, 20.0)], 4, 30)
    assert plan.touched == [(1, 14), (9, 9), (9, 25)]
    assert plan.bounds == [0, 9, 9, 30]
    # rank 0 touches rows 9..13 owned by rank 2 and gets nothing to own from it
    assert plan.to_owner(0) == [(2, (9, 14), None)]
    assert plan.to_owner(2) == [(0, None, (9, 14))]
    assert plan.shared_rows(0) == [(9, 14)]
    assert plan.pairwise() and plan.overlaps(0) == [(2, (9, 14))]
    thin = RowPlan.from_scan_rows([(0.3, 11.9), (12.2, 14.7), (15.1, 30.5)], 6, 40)
    assert not thin.pairwise()  # rows 15..18 lie under three stripes


def _batch_order(rank, world):
    import tike_b200.random
    from tike_b200.communicators import Comm
    from tike_b200.ptycho.solvers._common import draw_sequence, peek_sequence
    tike_b200.random.randomizer_np = np.random.default_rng(100 + rank)  # unsynchronised ranks
    comm = Comm()
    s1, n1 = draw_sequence(6, False, comm)
    p1 = peek_sequence(6, False, comm, n1)
    s2, _ = draw_sequence(6, False, comm)
    return s1, n1, p1, s2


def test_batch_order_and_its_prediction_follow_rank_0():
    """Every rank visits the batches in rank 0's order (ptycho.py broadcasts it)
    and predicts rank 0's next order for the data-stream prefetch."""
    res = _spawn(_batch_order)
    want = np.random.default_rng(100)
    first = [int(x) for x in want.permutation(6)]
    second = [int(x) for x in want.permutation(6)]
    for s1, n1, p1, s2 in res:
        assert s1 == first and n1 == second and p1 == second and s2 == second
