"""The CPU oracle against the reference's known answers and golden vectors.

Golden vectors in tests/golden/*.npz were produced by the unmodified reference
(tests/golden/make_golden.py).  Tolerances: float32 round-off only (1e-5
relative, L2) because both sides are NumPy/scipy.fft on the CPU.
"""
import numpy as np
import pytest

from oracle import ptycho_np as onp
from oracle import ref_shim
from conftest import load_golden, rel_err

TOL = 1e-5


def test_patch_fwd_known_answers():
    """Restates tests/operators/test_patch.py:64-133 (slices are the truth)."""
    size, win = 256, 8
    rng = np.random.default_rng(0)
    fov = (rng.random((size, size)) - 0.5 + 1j *
           (rng.random((size, size)) - 0.5)).astype(np.complex64)
    sub = 0.12346789
    c = size // 2 - win // 2
    positions = np.array([[0, 0], [0, size - win], [size - win, 0],
                          [size - win, size - win], [c, c], [sub, 3]],
                         dtype=np.float32)
    truth = np.stack([
        fov[:win, :win], fov[:win, -win:], fov[-win:, :win], fov[-win:, -win:],
        fov[c:c + win, c:c + win],
        (1.0 - sub) * fov[0:win, 3:3 + win] + sub * fov[1:1 + win, 3:3 + win],
    ])
    patches = onp.patch_fwd(fov, positions, win)
    np.testing.assert_allclose(patches, truth, atol=1e-6)


def test_patch_adj_known_answers():
    """Restates tests/operators/test_patch.py:136-206."""
    size, win = 8, 2
    positions = np.array([[0, 0], [0, size - win], [size - win, 0],
                          [size - win, size - win], [3, 3], [0.123, 3],
                          [3, 0.123], [5.5, 3.5]], dtype=np.float32)
    fov = np.zeros((size, size), dtype=np.complex64)
    for (y, x, w) in [(0, 0, 1), (0, size - win, 1), (size - win, 0, 1),
                      (size - win, size - win, 1), (3, 3, 1),
                      (0, 3, 1 - 0.123), (1, 3, 0.123), (3, 0, 1 - 0.123),
                      (3, 1, 0.123), (5, 3, .25), (6, 3, .25), (5, 4, .25),
                      (6, 4, .25)]:
        fov[y:y + win, x:x + win] += w
    out = onp.patch_adj(positions, np.ones((8, win, win), np.complex64),
                        np.zeros((size, size), np.complex64), win)
    np.testing.assert_allclose(out, fov, atol=1e-6)


def test_simulate_golden():
    """tests/ptycho/test_ptycho.py:191-203 with ptycho_setup.pickle.lzma."""
    g = load_golden('ptycho_setup')
    sim = onp.simulate(32, g['probe'], g['scan'], g['psi'])
    np.testing.assert_allclose(np.sqrt(sim), np.sqrt(g['data']), atol=1e-6)


@pytest.mark.parametrize('tag', [
    'rpie_batch_a', 'rpie_batch_pad', 'rpie_batch_poisson',
    'rpie_batch_poisson_dom', 'rpie_batch_eigen'
])
def test_rpie_batch_golden(tag):
    g = load_golden(tag)
    ew = g['eigen_weights'] if g['eigen_weights'].size else None
    costs, psi_num, probe_num, ew_out = onp.rpie_batch(
        g['data'], g['scan'], g['psi'], g['probe'], g['mask'],
        eigen_weights=ew, noise_model=str(g['noise_model']),
        unmeasured_scaling=float(g['scaling']), usemodes=str(g['usemodes']))
    assert rel_err(costs, g['costs']) < TOL
    assert rel_err(psi_num, g['psi_num']) < TOL
    assert rel_err(probe_num, g['probe_num']) < TOL
    pp = onp.psi_preconditioner(g['psi'], g['probe'], g['scan'])
    qp = onp.probe_preconditioner(g['psi'], g['probe'], g['scan'])
    assert rel_err(pp, g['psi_precond']) < TOL
    assert rel_err(qp, g['probe_precond']) < TOL
    psi_new, probe_new = onp.rpie_update(g['psi'], g['probe'], psi_num,
                                         probe_num, pp, qp, float(g['alpha']))
    assert rel_err(psi_new, g['psi_new']) < TOL
    assert rel_err(probe_new, g['probe_new']) < TOL
    if ew is not None:
        assert rel_err(ew_out, g['eigen_weights_out']) < TOL
    far = onp.farplane(g['psi'], g['scan'], g['probe'], int(g['det']))
    assert rel_err(onp.intensity(far), g['intensity']) < TOL


@pytest.mark.parametrize('tag', ['lstsq_batch_a', 'lstsq_batch_pad', 'lstsq_batch_poisson'])
def test_lstsq_batch_golden(tag):
    g = load_golden(tag)
    r = onp.lstsq_batch(g['data'], g['scan'], g['psi'], g['probe'], g['mask'],
                        g['psi_precond'], int(g['num_batch']),
                        recover_positions=True, noise_model=str(g['noise_model']))
    assert rel_err(r['chi'], g['chi']) < TOL
    assert rel_err(r['object_upd_sum'], g['obj_sum']) < TOL
    assert rel_err(r['m_probe_update'], g['m_probe_update']) < TOL
    assert rel_err(r['costs'], g['costs']) < TOL
    assert rel_err(r['patches'], g['patches']) < TOL
    assert rel_err(r['pos_num'], g['pos_num']) < TOL
    assert rel_err(r['pos_den'], g['pos_den']) < TOL
    assert rel_err(r['object_update_precond'], g['precond']) < TOL
    assert rel_err(r['beta_object'], g['beta_object']) < 1e-4
    assert rel_err(r['beta_probe'], g['beta_probe']) < 1e-4


@pytest.mark.skipif(not ref_shim.reference_available(),
                    reason='reference checkout not present on this box')
def test_reference_kats_pass_on_oracle_patch():
    """The reference's own KAT functions, executed on the oracle's patch."""
    import importlib
    import sys
    import types
    ref_shim.load_reference()
    pkg = types.ModuleType('reftests')
    pkg.__path__ = ['/root/reference/tests']
    sys.modules['reftests'] = pkg
    tp = importlib.import_module('reftests.operators.test_patch')
    tp.test_patch_correctness()
    tp.test_patch_correctness_adjoint()


def test_torch_standin_matches_oracle():
    """baseline/torch_standin.py (the GPU stand-in for the reference's CuPy op
    sequence that bench.py times) computes what the oracle computes."""
    import torch
    from baseline import torch_standin as ts
    from oracle import ptycho_np as onp
    from tike_b200 import synthetic
    psi, probe, scan = synthetic.make_problem(70, 16, 2, 60, 70, seed=3)
    data = onp.simulate(16, probe, scan, psi)
    rng = np.random.default_rng(0)
    psi2 = (psi * (1 + 0.1 * rng.standard_normal(psi.shape))).astype(np.complex64)
    c, pn, qn, _ = onp.rpie_batch(data, scan, psi2, probe, np.ones((16, 16), bool))
    c2, pn2, qn2 = ts.rpie_batch(torch.as_tensor(data), torch.as_tensor(scan),
                                 torch.as_tensor(psi2[0]), torch.as_tensor(probe[0, 0]))

    def rel(a, b):
        return np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(c2.numpy(), c) < 1e-5
    assert rel(pn2.numpy(), pn[0]) < 1e-5
    assert rel(qn2.numpy(), qn[0, 0, 0]) < 1e-5


def test_stripes_golden_vs_oracle_and_host_exchange():
    """The reference's 2-worker run (stripes_rpie golden) is reproduced by the
    oracle epoch per stripe + the product's host-side exchange helpers
    (swap_edges_pair, stitch_stripes): probe mean reaches worker 0 only (F11),
    halo blend of width N-1 at stripe_start[1], stitch at stripe_start + N//2."""
    import torch
    from oracle import ptycho_np as onp
    from tike_b200 import synthetic
    from tike_b200.communicators.comm import stitch_stripes, swap_edges_pair
    g = load_golden('stripes_rpie')
    det, N, M, P, H, W, seed = (int(g[k]) for k in ('det', 'N', 'M', 'P', 'H', 'W', 'seed'))
    psi_t, probe, scan = synthetic.make_problem(P, N, M, H, W, seed)
    data = onp.simulate(det, probe, scan, psi_t)
    order = [g['order0'], g['order1']]
    start = [int(s) for s in g['stripe_start']]
    mask = np.ones((det, det), bool)
    psi = [np.full_like(psi_t, 0.5 + 0j) for _ in range(2)]
    nd = ni = 0.0
    for o in order:  # ptycho.py:873-972 initial probe rescale over all workers
        inten = onp.intensity(onp.farplane(psi[0], scan[o], probe, det))
        nd += float(np.sum(data[o], dtype=np.float64))
        ni += float(np.sum(inten, dtype=np.float64))
    pr = [(probe * np.float32(np.sqrt(nd) / np.sqrt(ni))).astype(np.complex64)
          for _ in range(2)]
    lo, hi = start[1], start[1] + N - 1
    for ep in range(int(g['num_iter'])):
        costs = []
        pre = [onp.psi_preconditioner(psi[r], pr[r], scan[order[r]]) for r in range(2)]
        for r in range(2):
            o = order[r]
            psi[r], pr[r], _, c = onp.rpie_epoch(
                data[o], scan[o], psi[r], pr[r], mask, [np.arange(len(o))], [0],
                alpha=float(g['alpha']))
            costs.append(c)
        np.testing.assert_allclose(costs, g['costs'][ep], rtol=2e-4)
        pr[0] = ((pr[0] + pr[1]) / 2).astype(np.complex64)
        a = torch.from_numpy(psi[0][..., lo:hi, :].copy())
        b = torch.from_numpy(psi[1][..., lo:hi, :].copy())
        psi[0][..., lo:hi, :] = swap_edges_pair(a, b, N - 1, True).numpy()
        psi[1][..., lo:hi, :] = swap_edges_pair(a, b, N - 1, False).numpy()
        if (ep + 1) % 10 == 0:
            # remove_object_ambiguity (object.py:324-335) every rescale_period
            # epochs, per worker, with that epoch's object preconditioner
            for r in range(2):
                w = pre[r].real / onp.mnorm(pre[r].real)
                norm = 2 * np.sqrt(np.mean(np.square(np.abs(psi[r])) * w))
                psi[r] = (psi[r] / norm).astype(np.complex64)
                pr[r] = (pr[r] * norm).astype(np.complex64)
    joined = stitch_stripes([p.copy() for p in psi], start, N)
    assert np.linalg.norm(joined - g['psi']) / np.linalg.norm(g['psi']) < 1e-4
    assert np.linalg.norm(pr[0] - g['probe']) / np.linalg.norm(g['probe']) < 1e-4


def test_multislice_oracle_matches_reference():
    """Multislice (D = 3) forward model, rPIE gradients and preconditioners of
    the oracle against the reference (rpie_batch_ms golden)."""
    g = load_golden('rpie_batch_ms')
    N = int(g['N'])
    h = onp.fresnel_propagator(N, tuple(g['fov']), float(g['distance']), float(g['wavelength']))
    assert rel_err(h, g['propagator']) < 1e-6
    far = onp.multislice_farplane(g['psi'], g['scan'], g['probe'], h)
    assert rel_err(far, g['farplane']) < TOL
    mask = np.ones((N, N), bool)
    costs, psi_num, probe_num, _ = onp.rpie_batch_multislice(
        g['data'], g['scan'], g['psi'], g['probe'], mask, h)
    assert rel_err(costs, g['costs']) < TOL
    assert rel_err(psi_num, g['psi_num']) < TOL
    assert rel_err(probe_num, g['probe_num']) < TOL
    assert rel_err(onp.psi_preconditioner_multislice(g['psi'], g['probe'], g['scan'], h),
                   g['psi_precond']) < TOL
    assert rel_err(onp.probe_preconditioner_multislice(g['psi'], g['probe'], g['scan']),
                   g['probe_precond']) < TOL
    # the product's host-side Fresnel kernel is the same formula
    from tike_b200.kernels import fresnel_propagator
    assert rel_err(fresnel_propagator(N, tuple(g['fov']), float(g['distance']),
                                      float(g['wavelength'])), g['propagator']) < 1e-6


def test_probe_helpers_against_reference_fixtures():
    """The reference's known-answer fixtures for the per-epoch probe helpers
    (tests/ptycho/test_probe.py:137-178), converted to probe_fixtures.npz:
    orthogonalize_eig against ortho-out.mat (magnitudes, rtol 1e-4 as in the
    reference test — phases may differ by pi) and against the reference code's
    own output; add_modes_cartesian_hermite against hermite.mat."""
    import torch
    from tike_b200.ptycho import probe as P
    g = load_golden('probe_fixtures')
    got, power = P.orthogonalize_eig(torch.as_tensor(g['ortho_in']))
    got = got.numpy()
    np.testing.assert_allclose(np.abs(got), np.abs(g['ortho_out']), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.abs(got), np.abs(g['ortho_ref']), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.asarray(power), g['ortho_ref_power'], rtol=1e-4)
    her = P.add_modes_cartesian_hermite(g['hermite_in'][None, None, ...], 12)
    np.testing.assert_allclose(her, g['hermite_out'][None, ...], rtol=1e-4, atol=1e-6)
    assert rel_err(her, g['hermite_ref']) < 1e-5
