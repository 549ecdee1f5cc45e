"""Multi-GPU (one process per GPU, NCCL) parity: G ranks with all-reduced
numerators must equal ONE rank processing the union batches (SURVEY §8e mode
B).  Needs >= 2 GPUs; skipped otherwise (the driver's 1-GPU run skips it)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _problem():
    from oracle import ptycho_np as onp
    from tike_b200 import synthetic
    det = N = 32
    psi_t, probe, scan = synthetic.make_problem(160, N, 2, 120, 128, seed=21)
    data = onp.simulate(det, probe, scan, psi_t)
    return data, np.full_like(psi_t, 0.5 + 0j), probe, scan, det


def _params(tp, probe, psi0, scan, det, algo):
    alg = (tp.RpieOptions(num_batch=3, num_iter=6, alpha=0.3) if algo == 'rpie'
           else tp.LstsqOptions(num_batch=3, num_iter=6))
    return tp.PtychoParameters(
        probe=probe.copy(), psi=psi0.copy(), scan=scan.copy(), algorithm_options=alg,
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool)),
        probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions())


def _worker(rank, world, port, algo, mode, out):
    import torch.distributed as dist
    import tike_b200.ptycho as tp
    import tike_b200.random
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        data, psi0, probe, scan, det = _problem()
        tike_b200.random.randomizer_np = np.random.default_rng(5)
        np.random.seed(5)
        with tp.Reconstruction(data, _params(tp, probe, psi0, scan, det, algo),
                               multi_gpu_mode=mode) as ctx:
            split = (ctx.order, None, ctx.stripe_start)
            batches = ctx.comm.allgather_object([b.tolist() for b in ctx.batches])
            ctx.iterate(6)
            r = ctx.get_result()
            # the replicas of object and probe must be bit-identical on all ranks
            digests = ctx.comm.allgather_object(
                (r.psi.tobytes(), r.probe.tobytes()))
            out[f'replicas_identical_{rank}'] = all(d == digests[0] for d in digests)
            if mode == 'halo':
                out['plan'] = (ctx.comm.plan.touched, ctx.comm.plan.bounds,
                               list(ctx.comm.batch_cuts))
        if rank == 0:
            out['costs'] = [c[0] for c in r.algorithm_options.costs]
            out['psi'] = r.psi
            out['probe'] = r.probe
            out['order'] = [o.tolist() for o in split[0]]
            out['batches'] = batches
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['halo', 'allreduce'])
@pytest.mark.parametrize('algo', ['rpie', 'lstsq_grad'])
def test_two_ranks_equal_union_batches(algo, mode):
    """Both multi-GPU data planes -- 'halo' (row-bounded exchange overlapped
    with the batch kernel, every rank owning a row range) and 'allreduce'
    (whole-object NCCL all-reduce) -- against ONE rank fed the union batches."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    import tike_b200.ptycho as tp
    import tike_b200.random
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), algo, mode, out), nprocs=2, join=True)
    assert out['replicas_identical_0'] and out['replicas_identical_1']
    if mode == 'halo':
        touched, bounds, cuts = out['plan']
        print('row plan', touched, bounds, 'rank-0 cuts', cuts)
        assert bounds[0] == 0 and bounds[-1] == 120 and len(bounds) == 3

    # single rank over the union batches: batch n = rank0.batch n + rank1.batch n
    data, psi0, probe, scan, det = _problem()
    order = [np.array(o) for o in out['order']]
    union, ranges, lo = [], [], 0
    for n in range(3):
        idx = np.concatenate([order[g][np.array(out['batches'][g][n])] for g in range(2)])
        union.append(idx)
        ranges.append(np.arange(lo, lo + len(idx)))
        lo += len(idx)
    split = ([np.concatenate(union)], [ranges], [0])
    tike_b200.random.randomizer_np = np.random.default_rng(5)
    np.random.seed(5)
    with tp.Reconstruction(data, _params(tp, probe, psi0, scan, det, algo), split=split) as ctx:
        ctx.iterate(6)
        r = ctx.get_result()
    costs = np.array([c[0] for c in r.algorithm_options.costs])
    multi = np.array(out['costs'])
    rel = np.abs(costs - multi) / np.abs(costs)
    print(algo, 'multi-GPU vs union-batch cost rel err', rel)
    assert rel.max() < 1e-3
    def rel_err(a, b):
        return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
    assert rel_err(out['psi'], r.psi) < 1e-3
    assert rel_err(out['probe'], r.probe) < 1e-3


def _stripes_worker(rank, world, port, tag, out):
    import torch.distributed as dist
    import tike_b200.ptycho as tp
    import tike_b200.random
    from oracle import ptycho_np as onp
    from tike_b200 import synthetic
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        g = np.load(os.path.join(os.path.dirname(__file__), 'golden', tag + '.npz'))
        det, N, M, P, H, W, seed = (int(g[k]) for k in ('det', 'N', 'M', 'P', 'H', 'W', 'seed'))
        psi_t, probe, scan = synthetic.make_problem(P, N, M, H, W, seed)
        data = onp.simulate(det, probe, scan, psi_t)
        algo = str(g['algo'])
        nb, it = int(g['num_batch']), int(g['num_iter'])
        alg = (tp.RpieOptions(num_batch=nb, num_iter=it, alpha=float(g['alpha']))
               if algo == 'rpie' else tp.LstsqOptions(num_batch=nb, num_iter=it))
        params = tp.PtychoParameters(
            probe=probe.copy(), psi=np.full_like(psi_t, 0.5 + 0j), scan=scan.copy(),
            algorithm_options=alg,
            exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool)),
            probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions())
        tike_b200.random.randomizer_np = np.random.default_rng(seed)
        np.random.seed(seed)
        with tp.Reconstruction(data, params, multi_gpu_mode='stripes') as ctx:
            order, start = ctx.cluster_order, ctx.stripe_start
            ctx.iterate(it)
            r = ctx.get_result()
        if rank == 0:
            out['costs'] = np.array(r.algorithm_options.costs)
            out['psi'], out['probe'] = r.psi, r.probe
            out['order'] = [np.asarray(o) for o in order]
            out['stripe_start'] = list(start)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('tag', ['stripes_rpie', 'stripes_lstsq'])
def test_stripes_mode_matches_reference_two_workers(tag):
    """multi_gpu_mode='stripes' (independent stripes + probe mean on worker 0 +
    halo blend + stitch) against the reference run with num_gpu=2."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', tag + '.npz'))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_stripes_worker, args=(2, _free_port(), tag, out), nprocs=2, join=True)
    for i in range(2):  # stripe assignment is integer work: bit-exact
        np.testing.assert_array_equal(out['order'][i], g[f'order{i}'])
    np.testing.assert_array_equal(out['stripe_start'], g['stripe_start'])
    rel = np.abs(out['costs'] - g['costs']) / np.abs(g['costs'])
    print(tag, 'stripes-mode cost rel err per epoch', rel.max(axis=1))
    assert rel.max() < 1e-3

    def rel_err(a, b):
        return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
    assert rel_err(out['psi'], g['psi']) < 2e-3
    assert rel_err(out['probe'], g['probe']) < 2e-3


def _options_worker(rank, world, port, out):
    import torch.distributed as dist
    import tike_b200.ptycho as tp
    import tike_b200.random
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        data, psi0, probe, scan, det = _problem()
        tike_b200.random.randomizer_np = np.random.default_rng(5)
        np.random.seed(5)
        results = {}
        for algo in ('rpie', 'lstsq_grad'):
            eigen_probe, weights = tp.probe.init_varying_probe(scan, probe, num_eigen_probes=1,
                                                               probes_with_modes=1)
            alg = (tp.RpieOptions(num_batch=2, num_iter=3, alpha=0.3) if algo == 'rpie'
                   else tp.LstsqOptions(num_batch=2, num_iter=3))
            params = tp.PtychoParameters(
                probe=probe.copy(), psi=psi0.copy(), scan=scan.copy(), algorithm_options=alg,
                eigen_probe=eigen_probe, eigen_weights=weights,
                exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool)),
                probe_options=tp.ProbeOptions(force_orthogonality=True),
                object_options=tp.ObjectOptions(use_adaptive_moment=True),
                position_options=(tp.PositionOptions(initial_scan=scan.copy(),
                                                     update_magnitude_limit=0.5,
                                                     use_position_regularization=True)
                                  if algo == 'lstsq_grad' else None))
            r = tp.reconstruct(data, params)
            results[algo] = (np.array(r.algorithm_options.costs), r.scan.shape,
                             r.eigen_weights.shape, bool(np.all(np.isfinite(r.psi))))
        if rank == 0:
            out.update(results)
    finally:
        dist.destroy_process_group()


def test_two_ranks_with_varying_probe_positions_and_momentum():
    """Optional features through the multi-rank driver: eigen weights, position
    correction with affine regularisation, adaptive moment (results are
    gathered in the caller's position order; costs finite and decreasing)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_options_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for algo in ('rpie', 'lstsq_grad'):
        costs, scan_shape, w_shape, finite = out[algo]
        assert finite and np.all(np.isfinite(costs))
        assert costs.shape == (3, 2) and costs[-1].mean() < costs[0].mean()
        assert scan_shape == (160, 2) and w_shape[0] == 160
