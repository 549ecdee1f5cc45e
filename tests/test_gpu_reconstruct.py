"""End-to-end parity of tike_b200.ptycho.reconstruct with the reference.

The golden trajectories in tests/golden/traj_*.npz were produced by the
unmodified reference (tests/golden/make_golden.py) on the same seeded
synthetic inputs, which are rebuilt here from tike_b200.synthetic and the
oracle's forward model.  Bars (BASELINE.json north_star): batch assignment
bit-exact; cost trajectory within 1e-3 relative after 50 epochs.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _setup(g):
    from oracle import ptycho_np as onp
    from tike_b200 import synthetic
    position = bool(g['position'])
    det, N, M, P, H, W, seed = (int(g[k]) for k in ('det', 'N', 'M', 'P', 'H', 'W', 'seed'))
    psi_true, probe, scan = synthetic.make_problem(
        P, N, M, H, W, seed, margin=6.0 if position else 0.0)
    data = onp.simulate(det, probe, scan, psi_true)
    psi0 = np.full_like(psi_true, 0.5 + 0j)
    rng = np.random.default_rng(seed + 30)
    scan0 = scan
    if position:
        scan0 = (scan + rng.uniform(-0.6, 0.6, scan.shape)).astype(np.float32)
    return data, psi0, probe, scan0, det, seed


def _run(tag):
    import tike_b200.ptycho as tp
    import tike_b200.random
    g = load_golden(tag)
    data, psi0, probe, scan0, det, seed = _setup(g)
    algo = str(g['algo'])
    common = dict(num_batch=int(g['num_batch']), num_iter=int(g['num_iter']),
                  batch_method=str(g['batch_method']))
    alg = (tp.RpieOptions(alpha=float(g['alpha']), **common) if algo == 'rpie'
           else tp.LstsqOptions(**common))
    params = tp.PtychoParameters(
        probe=probe.copy(), psi=psi0, scan=scan0.copy(), algorithm_options=alg,
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool)),
        probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions(),
        position_options=tp.PositionOptions(initial_scan=scan0.copy(),
                                            update_magnitude_limit=1.0)
        if bool(g['position']) else None)
    tike_b200.random.randomizer_np = np.random.default_rng(seed)
    np.random.seed(seed)
    with tp.Reconstruction(data, params) as ctx:
        order = ctx.order[0]
        sizes = np.array([len(b) for b in ctx.batches])
        ctx.iterate(alg.num_iter)
        result = ctx.get_result()
    return g, order, sizes, result


@pytest.mark.parametrize('tag', ['traj_rpie', 'traj_rpie_compact', 'traj_lstsq'])
def test_trajectory_matches_reference(tag):
    g, order, sizes, result = _run(tag)
    assert np.array_equal(order, g['order']), 'batch assignment must be bit-exact'
    assert np.array_equal(sizes, g['batch_sizes'])
    costs = np.array([c[0] for c in result.algorithm_options.costs])
    ref = g['costs']
    assert len(costs) == len(ref)
    rel = np.abs(costs - ref) / np.abs(ref)
    print(tag, 'cost rel err: first', rel[0], 'max', rel.max(), 'last', rel[-1])
    assert rel[-1] < 1e-3, f'final cost off by {rel[-1]:.2e}'
    assert rel.max() < 5e-3
    assert rel_err(result.psi, g['psi']) < 5e-3
    assert rel_err(result.probe, g['probe']) < 5e-3
    assert rel_err(result.probe_options.power[-1], g['probe_power']) < 5e-3


def test_position_correction_trajectory():
    g, order, sizes, result = _run('traj_lstsq_pos')
    assert np.array_equal(order, g['order'])
    costs = np.array([c[0] for c in result.algorithm_options.costs])
    rel = np.abs(costs - g['costs']) / np.abs(g['costs'])
    print('position cost rel err max', rel.max(), 'last', rel[-1])
    assert rel[-1] < 1e-3
    assert np.abs(result.scan - g['scan']).max() < 2e-2


def test_reconstruct_twice_continues():
    """templates.py:115-125: feeding the result back in continues the run."""
    import tike_b200.ptycho as tp
    g = load_golden('traj_rpie')
    data, psi0, probe, scan0, det, seed = _setup(g)
    params = tp.PtychoParameters(
        probe=probe.copy(), psi=psi0, scan=scan0.copy(),
        algorithm_options=tp.RpieOptions(num_batch=3, num_iter=3, alpha=0.2),
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool)),
        probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions())
    for _ in range(2):
        params = tp.reconstruct(data, params)
    costs = [c[0] for c in params.algorithm_options.costs]
    assert len(costs) == 6 and costs[-1] < costs[0]
    assert isinstance(params.psi, np.ndarray) and params.psi.dtype == np.complex64


def test_simulate_matches_golden():
    import tike_b200.ptycho as tp
    g = load_golden('ptycho_setup')
    out = tp.simulate(32, g['probe'], g['scan'], g['psi'])
    np.testing.assert_allclose(np.sqrt(out), np.sqrt(g['data']), atol=1e-6)


def test_shape_errors_match_reference():
    import tike_b200.ptycho as tp
    probe = np.ones((1, 1, 1, 16, 16), np.complex64)
    psi = np.ones((1, 64, 64), np.complex64)
    scan = np.full((4, 2), 5, np.float32)
    with pytest.raises(ValueError):
        tp.PtychoParameters(probe=probe, psi=psi[0], scan=scan)  # psi must be 3-D
    with pytest.raises(ValueError):
        tp.PtychoParameters(probe=probe, psi=psi, scan=scan * 100)  # outside FOV
    p = tp.PtychoParameters(probe=probe, psi=psi, scan=scan)
    with pytest.raises(ValueError):
        tp.Reconstruction(np.ones((3, 16, 16), np.float32), p)  # frames != positions
