"""CPU-only tests: C-ABI library surface, host-side logic, options API."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_err


def test_library_exports_every_declared_symbol():
    """Every function declared in include/tike_b200.h is exported by the
    shared library and bound by the ctypes layer (no compute calls here)."""
    from tike_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'tike_b200.h')).read()
    declared = set(re.findall(r'\b(tb_[a-z0-9_]+)\s*\(', header))
    declared -= {'tb_stream_t'}
    assert declared, 'no declarations parsed'
    if not os.path.exists(_lib.LIB_PATH):
        from tike_b200 import build
        build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(handle, name)]
    assert not missing, f'library lacks {missing}'
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib().tb_version() >= 100


def test_struct_layout_matches_header():
    """ctypes structs must have the C layout (spot-check sizes/offsets)."""
    from tike_b200 import _lib
    assert _lib.tb_batch.psi.offset == 0
    assert _lib.tb_batch.scan.offset == 16
    assert _lib.tb_batch.npos.offset == 24
    assert _lib.tb_batch.probe.offset == 32
    assert _lib.tb_batch.eigen_probe.offset == 56
    assert _lib.tb_batch.eigen_weights.offset == 72
    assert ctypes.sizeof(_lib.tb_batch) == 96
    assert _lib.tb_rpie_args.data.offset == 96


def test_compute_path_refuses_host_arrays():
    from tike_b200 import _lib
    with pytest.raises(TypeError):
        _lib.dev_ptr(np.zeros(4, np.float32), '<f4', 'x')


@pytest.mark.parametrize('method', ['wobbly_center', 'compact',
                                    'wobbly_center_random_bootstrap'])
@pytest.mark.parametrize('nworker', [1, 2, 3])
def test_cluster_bit_exact_with_reference(method, nworker):
    """Orders / batch sizes / stripe starts equal the reference's
    cluster.by_scan_stripes_contiguous output index for index."""
    from tike_b200 import cluster
    g = load_golden('cluster')
    np.random.seed(int(g['seed']))
    order, batches, start = cluster.by_scan_stripes_contiguous(
        g['scan'], nworker, method, 4)
    for i in range(nworker):
        assert np.array_equal(order[i], g[f'{method}_{nworker}_{i}_order'])
        assert np.array_equal([len(b) for b in batches[i]],
                              g[f'{method}_{nworker}_{i}_sizes'])
        # batches are contiguous ranges covering the worker's positions
        assert np.array_equal(np.concatenate(batches[i]), np.arange(len(order[i])))
    assert np.array_equal(start, g[f'{method}_{nworker}_start'])


def test_cluster_complete_and_sized():
    """tests/test_random.py:12-228 properties: every index exactly once."""
    from tike_b200 import cluster
    rng = np.random.default_rng(0)
    pop = rng.random((101, 2)).astype(np.float32)
    for f in (cluster.wobbly_center, cluster.compact,
              cluster.wobbly_center_random_bootstrap):
        groups = f(pop, 5)
        allidx = np.sort(np.concatenate(groups))
        assert np.array_equal(allidx, np.arange(101))
        sizes = sorted(len(g) for g in groups)
        assert sizes[-1] - sizes[0] <= 1
    stripes = cluster.stripes_equal_count(pop, 3)
    assert np.array_equal(np.sort(np.concatenate(stripes)), np.arange(101))
    assert pop[stripes[0], 0].max() <= pop[stripes[1], 0].min()
    masks = cluster.by_scan_stripes(pop, 3)
    assert np.sum(masks) == 101


def test_options_api_and_validation():
    import tike_b200.ptycho as tp
    assert tp.RpieOptions().name == 'rpie' and tp.RpieOptions().num_batch == 5
    assert tp.RpieOptions().alpha == 0.05
    assert tp.LstsqOptions().name == 'lstsq_grad'
    assert tp.DmOptions().name == 'dm' and tp.DmOptions().num_batch == 1
    probe = np.ones((1, 1, 2, 16, 16), np.complex64)
    psi = np.ones((1, 64, 64), np.complex64)
    scan = np.full((4, 2), 5.5, np.float32)
    p = tp.PtychoParameters(probe=probe, psi=psi, scan=scan)
    assert p.exitwave_options.measured_pixels.shape == (16, 16)
    with pytest.raises(ValueError):
        tp.PtychoParameters(probe=probe[0], psi=psi, scan=scan)
    with pytest.raises(ValueError):
        tp.PtychoParameters(probe=probe, psi=psi, scan=scan[:, :1])
    with pytest.raises(ValueError):
        tp.PtychoParameters(probe=probe, psi=psi, scan=scan + 60)
    # split keeps dtype policy and subset
    q = tp.PtychoParameters.split(np.array([2, 0]), x=p)
    assert q.scan.shape == (2, 2) and q.psi.dtype == np.complex64
    po = tp.ProbeOptions(update_start=3, update_period=2)
    assert [po.recover_probe(e) for e in range(6)] == [False] * 4 + [True, False]


def test_affine_transform_roundtrip():
    from tike_b200.ptycho.position import AffineTransform
    t = AffineTransform(scale0=1.1, scale1=0.9, shear1=0.05, angle=0.1, t0=1, t1=-2)
    back = AffineTransform.fromarray(t.asarray3())
    np.testing.assert_allclose(back.astuple(), t.astuple(), atol=1e-5)
    np.testing.assert_allclose(AffineTransform.frombuffer(t.asbuffer()).astuple(),
                               t.astuple())


def test_fit_line():
    """tests/test_opt.py:20-25."""
    from tike_b200 import opt
    slope, intercept = opt.fit_line_least_squares(
        y=np.asarray([0, np.log(0.9573), np.log(0.8386)]), x=np.asarray([0, 1, 2]))
    assert abs(slope - (-0.08801)) < 1e-4 and abs(intercept - 0.014789) < 1e-3


def test_probe_helpers_host():
    from tike_b200.ptycho import probe as P
    rng = np.random.default_rng(0)
    base = (rng.random((1, 1, 1, 16, 16)) + 1j * rng.random((1, 1, 1, 16, 16))).astype(np.complex64)
    modes = P.add_modes_cartesian_hermite(base, 5)
    assert modes.shape == (1, 1, 5, 16, 16)
    gram = np.einsum('mij,nij->mn', modes[0, 0].conj(), modes[0, 0])
    np.testing.assert_allclose(gram, np.eye(5), atol=1e-4)
    modes = P.adjust_probe_power(modes)
    power = np.sum(np.abs(modes[0, 0])**2, axis=(-2, -1))
    np.testing.assert_allclose(power / power[0], (1.0 / np.arange(1, 6))**2, rtol=1e-4)
    np.random.seed(0)
    eig, w = P.init_varying_probe(np.zeros((7, 2), np.float32), modes, 3, 2)
    assert eig.shape == (1, 2, 2, 16, 16) and w.shape == (7, 3, 5)


def test_host_fft_unit_test_binary(tmp_path):
    """Compile the FFT building blocks for the host and check them against a
    naive DFT (radix butterflies, digit-reversed order, inverse)."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    exe = tmp_path / 'fft_host_test'
    src = os.path.join(ROOT, 'tests', 'csrc', 'fft_host_test.cu')
    subprocess.run([nvcc, '-std=c++17', '--expt-relaxed-constexpr', '-O2', '-o',
                    str(exe), src], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'PASS' in out.stdout


def test_native_cluster_loop_is_bit_exact():
    """The C growth loop (tb_cluster_grow) must reproduce the NumPy loop of
    wobbly_center index for index, including duplicate points (ties)."""
    from tike_b200 import cluster
    rng = np.random.default_rng(7)
    pops = [(rng.random((1500, 2)) * 3900 + 2).astype(np.float32),
            np.repeat((rng.random((700, 2)) * 100).astype(np.float32), 2, axis=0)]
    for pop, nc in zip(pops, (4, 3)):
        native = cluster.wobbly_center(pop, nc)
        saved = cluster._NATIVE_MIN_POINTS
        cluster._NATIVE_MIN_POINTS = 10**9
        try:
            ref = cluster.wobbly_center(pop, nc)
        finally:
            cluster._NATIVE_MIN_POINTS = saved
        assert all(np.array_equal(a, b) for a, b in zip(native, ref))


def test_num_gpu_without_distributed_launch_warns():
    """num_gpu > 1 in a single process (the reference's thread-pool model) is
    not silently accepted: one process per GPU is required."""
    import warnings
    import tike_b200.ptycho as tp
    data = np.ones((4, 16, 16), np.float32)
    params = tp.PtychoParameters(
        probe=np.ones((1, 1, 1, 16, 16), np.complex64),
        psi=np.ones((1, 40, 40), np.complex64),
        scan=np.full((4, 2), 5.0, np.float32),
        algorithm_options=tp.RpieOptions(),
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((16, 16), bool)),
        probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        tp.Reconstruction(data, params, num_gpu=2)
    assert any('torchrun' in str(x.message) for x in w)
    with pytest.raises(ValueError):
        tp.Reconstruction(data, params, multi_gpu_mode='ring')


def test_helper_api_of_the_reference_modules():
    """Small public helpers a tike user script may import: opt, linalg,
    exitwave step lengths, position.gaussian_gradient (checked against the
    oracle / NumPy where a counterpart exists)."""
    import torch
    from oracle import ptycho_np as onp
    from tike_b200 import linalg, opt
    from tike_b200.ptycho import exitwave, position

    class Options:
        convergence_window = 4
        costs = [[5.0], [4.0], [3.5], [3.6], [3.7], [3.8], [3.9], [4.0]]
    assert opt.is_converged(Options)
    Options.costs = [[float(10 - i)] for i in range(8)]
    assert not opt.is_converged(Options)
    assert [len(b) for b in opt.batch_indicies(10, 3, use_random=False)] == [4, 3, 3]
    g = np.ones(4, np.complex64)
    d, v, m = opt.adagrad(g)
    d, v, m = opt.adagrad(g, v, m)
    np.testing.assert_allclose(d.real, 1 / np.sqrt(2 + 1e-6), rtol=1e-6)

    rng = np.random.default_rng(0)
    a, b = rng.random((20, 3)), rng.random((20, 2))
    np.testing.assert_allclose(linalg.lstsq(a, b), np.linalg.lstsq(a, b, rcond=None)[0],
                               rtol=1e-8)
    x = rng.random((3, 50, 4)) + 1j * rng.random((3, 50, 4))
    S, U = linalg.pca_eig(x, 2)
    S2, U2 = linalg.pca_eig(torch.as_tensor(x), 2)
    np.testing.assert_allclose(S2.numpy(), S, rtol=1e-10)
    assert S.shape == (3, 2) and U.shape == (3, 4, 2) and np.all(S[:, 0] >= S[:, 1])

    w = (rng.random((3, 20, 24)) + 1j * rng.random((3, 20, 24))).astype(np.complex64)
    gy, gx = position.gaussian_gradient(w)
    oy, ox = onp.gaussian_gradient(w)
    ty, tx = position.gaussian_gradient(torch.as_tensor(w))
    assert rel_err(gy, oy) < 1e-6 and rel_err(gx, ox) < 1e-6
    assert rel_err(ty.numpy(), oy) < 1e-6 and rel_err(tx.numpy(), ox) < 1e-6

    F, S_, Wd = 4, 3, 16
    I_m = rng.random((F, Wd, Wd)).astype(np.float32) + 0.1
    ab = rng.random((F, 1, S_, Wd, Wd)).astype(np.float32)
    I_e = ab.sum(2)[:, 0]
    xi = (1 - I_m / I_e)[:, None, None]
    mask = np.ones((Wd, Wd), bool)
    mask[2:4] = False
    st = np.full((F, 1, S_, 1, 1), 0.5, np.float32)
    ref = onp._poisson_steps_all_modes(xi, ab, I_e, I_m, mask, st, 0.5)
    got = exitwave.poisson_steplength_all_modes(xi, ab, I_e, I_m, mask, st, 0.5)
    assert rel_err(got, ref) < 1e-6
    got_t = exitwave.poisson_steplength_all_modes(
        *(torch.as_tensor(z) for z in (xi, ab, I_e, I_m, mask, st)), 0.5)
    assert rel_err(got_t.numpy(), ref) < 1e-5
    dom = exitwave.poisson_steplength_dominant_mode(xi, I_e, I_m, mask, st, 0.5)
    assert dom.shape == st.shape and np.all(np.isfinite(dom))


def test_native_compact_sweep_is_bit_exact(monkeypatch):
    """cluster.compact with the swap sweep in libtikeb200 equals the pure
    NumPy loop (same labels), on a size where the native path is taken."""
    from tike_b200 import cluster, synthetic
    scan = synthetic.make_scan(1500, 600, 600, 32, seed=4)
    np.random.seed(3)
    native = [np.asarray(a) for a in cluster.compact(scan, 4)]
    monkeypatch.setattr(cluster, '_native_compact_sweep', lambda *a, **k: None)
    np.random.seed(3)
    pure = [np.asarray(a) for a in cluster.compact(scan, 4)]
    assert len(native) == len(pure)
    for a, b in zip(native, pure):
        np.testing.assert_array_equal(a, b)


def test_split_by_scan_known_answer():
    """tests/test_random.py:207-228 of the reference, restated."""
    from tike_b200 import cluster
    scan = np.moveaxis(np.mgrid[0:3, 0:3].reshape(2, -1), 0, -1)
    split = [scan[i] for i in cluster.by_scan_stripes(scan, 3, axis=0)]
    np.testing.assert_equal(split, [[[0, 0], [0, 1], [0, 2]],
                                    [[1, 0], [1, 1], [1, 2]],
                                    [[2, 0], [2, 1], [2, 2]]])
    split = [scan[i] for i in cluster.by_scan_stripes(scan, 3, axis=1)]
    np.testing.assert_equal(split, [[[0, 0], [1, 0], [2, 0]],
                                    [[0, 1], [1, 1], [2, 1]],
                                    [[0, 2], [1, 2], [2, 2]]])


def test_reference_position_and_probe_unit_tests_restated():
    """tests/ptycho/test_position.py:22-107 and tests/ptycho/test_probe.py:60-136
    of the reference, restated for this package: PositionOptions split/join,
    AffineTransform known answers, varying-probe shapes, probe-support bounds."""
    import torch
    import tike_b200.ptycho as tp
    from tike_b200.ptycho import probe as P
    rng = np.random.default_rng(0)
    # --- PositionOptions.split / join
    N, num_batch = 245, 11
    scan = rng.random((N, 2)).astype(np.float32)
    indices = rng.permutation(N)
    batches = np.array_split(indices, num_batch)
    reorder = np.argsort(np.concatenate(batches))
    opts = tp.PositionOptions(scan, use_adaptive_moment=True)
    joined = tp.PositionOptions.join([opts.split(b) for b in batches], reorder=reorder)
    np.testing.assert_array_equal(joined.initial_scan, opts.initial_scan)
    # --- AffineTransform known answers
    pts = np.array([[0, 0], [0, 1], [1, 0], [-1, -1]])
    np.testing.assert_allclose(tp.AffineTransform(t0=11, t1=-5)(pts),
                               [[11, -5], [11, -4], [12, -5], [10, -6]])
    np.testing.assert_allclose(tp.AffineTransform(scale0=11, scale1=0.5)(pts),
                               [[0, 0], [0, 0.5], [11, 0], [-11, -0.5]])
    # --- weighted linear fit recovers a composed transform
    truth = [3.4567, 5.4321, 0.9876, 1.2345, 2.3456, -4.5678]
    T = tp.AffineTransform(*truth)
    p0 = rng.random((213, 2)) - 0.5
    err = rng.normal(size=(213, 2), scale=0.01)
    p1 = T(p0) + err
    from tike_b200 import linalg
    fit = tp.AffineTransform.fromarray(
        linalg.lstsq(a=np.pad(p0, ((0, 0), (0, 1)), constant_values=1), b=p1,
                     weights=1 / (1 + np.square(err).sum(axis=-1))))
    np.testing.assert_allclose(fit.asarray3(), T.asarray3(), atol=0.02)
    # --- get_varying_probe / init_varying_probe shapes
    for p, e, s, vary in [(0, 0, 1, False), (0, 0, 7, False), (31, 0, 1, True),
                          (31, 0, 7, True), (31, 3, 1, True), (31, 3, 7, True)]:
        unique = P.get_varying_probe(
            torch.rand(1, 1, s, 16, 16, dtype=torch.float32).to(torch.complex64),
            torch.rand(1, e, s, 16, 16).to(torch.complex64) if e > 0 else None,
            torch.ones(p, e + 1, s) if vary else None)
        assert tuple(unique.shape) == (p if vary else 1, 1, s, 16, 16)
    for p, e, s, w, v in [(31, 0, 2, 16, 0), (31, 0, 2, 16, 1), (31, 1, 2, 16, 1),
                          (31, 1, 3, 16, 3), (31, 7, 3, 16, 3), (31, 7, 3, 16, 1)]:
        eigen_probe, weights = P.init_varying_probe(
            scan=rng.random((p, 2)), shared_probe=rng.random((1, 1, s, w, w)),
            num_eigen_probes=e, probes_with_modes=v)
        assert (eigen_probe is None) if e < 2 else eigen_probe.shape == (1, e - 1, v, w, w)
        assert (weights is None) if e < 1 else weights.shape == (p, e, s)
    # --- finite probe support penalty bounds
    penalty = P.finite_probe_support(torch.zeros((101, 101)), radius=0.5 * 0.7, degree=2.5,
                                     p=2.345)
    penalty = np.asarray(penalty)
    assert round(float(penalty.min()), 3) == 0.000 and round(float(penalty.max()), 3) == 2.345


def test_reference_ptycho_utils_known_answers():
    """tests/ptycho/test_ptycho.py:77-108 of the reference: probe.gaussian
    against its pickled fixture, check_allowed_positions, get_padded_object."""
    import tike_b200.ptycho as tp
    g = load_golden('ptycho_gaussian')
    np.testing.assert_array_equal(tp.probe.gaussian(15, rin=0.8, rout=1.0), g['weights'])
    psi = np.empty((1, 4, 9))
    probe = np.empty((8, 2, 2))
    tp.check_allowed_positions(np.array([[1, 1], [1, 6.9], [1.1, 1], [1.9, 5.5]]), psi,
                               probe.shape)
    for bad in np.array([[1, 7], [1, 0.9], [0.9, 1], [1, 0]]):
        with pytest.raises(ValueError):
            tp.check_allowed_positions(bad, psi, probe.shape)
    probe = np.empty((8, 3, 4))
    scan = (np.random.default_rng(0).random((15, 2)) * 100) - 50
    psi, scan = tp.object.get_padded_object(scan, probe)
    tp.check_allowed_positions(scan, psi, probe_shape=probe.shape)


def test_reference_linalg_unit_tests_restated(monkeypatch):
    """tests/test_linalg.py of the reference, restated (NumPy and torch)."""
    import torch
    from tike_b200 import linalg
    from tike_b200 import random as tb_random
    # seeded: an unseeded draw now and then gives a 3x3 system too
    # ill-conditioned for complex64 at rtol=1e-2
    monkeypatch.setattr(tb_random, 'randomizer_np', np.random.default_rng(7))
    a = tb_random.numpy_complex(5)
    np.testing.assert_allclose(np.sqrt(linalg.inner(a, a).real), np.linalg.norm(a), rtol=1e-6)
    np.testing.assert_allclose(linalg.norm(a), np.linalg.norm(a), rtol=1e-6)
    A = tb_random.numpy_complex(5, 1, 4, 3, 3)
    x = tb_random.numpy_complex(5, 1, 4, 3, 1)
    w = np.random.default_rng(0).random((5, 1, 4, 3)).astype(np.float32)
    np.testing.assert_allclose(linalg.lstsq(A, A @ x, weights=w), x, rtol=1e-2, atol=1e-4)
    np.testing.assert_allclose(
        linalg.lstsq(torch.as_tensor(A), torch.as_tensor(A @ x), weights=torch.as_tensor(w)).numpy(),
        x, rtol=1e-2, atol=1e-4)
    b = tb_random.numpy_complex(5)
    assert abs(linalg.inner(a - linalg.projection(a, b), b)) < 1e-6
    assert abs(linalg.inner(a, b - linalg.projection(b, a))) < 1e-6
    v = tb_random.numpy_complex(1, 4, 3, 3)
    with pytest.raises(ValueError):
        linalg.orthogonalize_gs(v, axis=(0, 1, 2, 3))
    assert linalg.orthogonalize_gs(v).shape == v.shape
    assert linalg.orthogonalize_gs(v, axis=(1, -1)).shape == v.shape
    u = linalg.orthogonalize_gs(v, axis=(-2, -1))
    for i in range(4):
        for j in range(i + 1, 4):
            assert np.all(np.abs(linalg.inner(u[:, i:i + 1], u[:, j:j + 1], axis=(-2, -1))) < 1e-6)


@pytest.mark.parametrize('name', ['_resize_fft', '_resize_spline', '_resize_linear',
                                  '_resize_cubic', '_resize_lanczos'])
def test_resample_functions(name):
    """tests/ptycho/test_multigrid.py:26-56 of the reference: every resampling
    function handles factors 0.25 ... 4 on a probe-shaped array; the shape
    scales, and the OpenCV interpolators keep a constant constant."""
    from tike_b200.ptycho.solvers import options
    function = getattr(options, name)
    probe = np.full((1, 1, 2, 32, 32), 1 + 0.5j, np.complex64)
    for f in (0.25, 0.5, 1.0, 2.0, 4.0):
        out = function(probe, f)
        assert out.shape == (1, 1, 2, int(32 * f), int(32 * f))
        assert np.all(np.isfinite(out))
        if name in ('_resize_linear', '_resize_cubic', '_resize_lanczos'):
            np.testing.assert_allclose(out, 1 + 0.5j, rtol=1e-4)


def test_band_order_is_a_permutation_sorted_by_band_then_column():
    """kernels.band_order: the visiting order handed to the preconditioner
    kernels (csrc/precond.cu) — every position once, bands of PRECOND_BAND
    rows ascending, columns ascending inside a band; negative and empty
    inputs included."""
    import torch
    from tike_b200 import kernels
    rng = np.random.default_rng(5)
    scan = np.stack([rng.uniform(-20, 500, 3000), rng.uniform(-20, 700, 3000)], 1).astype(np.float32)
    order = kernels.band_order(torch.as_tensor(scan)).numpy()
    assert order.dtype == np.int32
    assert sorted(order.tolist()) == list(range(len(scan)))
    f = np.floor(scan).astype(np.int64)
    band = f[:, 0] // kernels.PRECOND_BAND
    assert np.all(np.diff(band[order]) >= 0)
    for b in np.unique(band):
        cols = f[order][band[order] == b, 1]
        assert np.all(np.diff(cols) >= 0)
    assert kernels.band_order(torch.zeros((0, 2))).shape == (0,)


@pytest.mark.parametrize('noise,outliers', [(0.3, False), (3.0, True)])
def test_native_ransac_inlier_pass_is_bit_exact(monkeypatch, noise, outliers):
    """tb_affine_inliers (csrc/cluster_host.cu) against the NumPy expressions
    it replaces in estimate_global_transformation_ransac (position.py:277-327
    of the reference): same transform, same fitness, for the all-inlier case
    and for a partial consensus."""
    import tike_b200.ptycho.position as pos
    from tike_b200 import random as tb_random
    rng = np.random.default_rng(0)
    P = 20000
    scan0 = rng.uniform(0, 4000, (P, 2)).astype(np.float32)
    moved = (scan0 * 1.001 + rng.normal(0, noise, (P, 2))).astype(np.float32)
    if outliers:
        moved[::7] += 100
    results = []
    native_pass = pos._native_inliers()
    assert native_pass is not None
    for native in (True, False):
        monkeypatch.setattr(tb_random, 'randomizer_np', np.random.default_rng(3))
        if not native:
            monkeypatch.setattr(pos, '_native_inliers', lambda: None)
        t, fitness = pos.estimate_global_transformation_ransac(scan0, moved, max_error=32)
        results.append((t.astuple(), fitness))
    assert results[0] == results[1]
    assert np.isfinite(results[0][1])
    # the mask itself, on a transform that splits the points
    x0, y0 = (np.ascontiguousarray(scan0[:, i], dtype=np.float64) for i in (0, 1))
    x1, y1 = (np.ascontiguousarray(moved[:, i], dtype=np.float64) for i in (0, 1))
    t = pos.AffineTransform(scale0=1.002, scale1=0.999, shear1=0.001, angle=0.0005, t0=0.5, t1=-0.25)
    mask = np.empty(P, dtype=np.uint8)
    count = native_pass(x0, y0, x1, y1, t, 6.0, mask)
    m = t.asarray().astype(np.float64)
    rx = x0 * m[0, 0] + y0 * m[1, 0] + t.t0 - x1
    ry = x0 * m[0, 1] + y0 * m[1, 1] + t.t1 - y1
    want = (rx * rx + ry * ry) <= 36.0
    assert 0 < want.sum() < P
    np.testing.assert_array_equal(mask.view(np.bool_), want)
    assert count == int(want.sum())


def test_band_sort_batches_keeps_the_partition():
    """cluster.band_sort_batches: same members in every batch, same ranges,
    every position once; inside a batch bands of BAND_ROWS rows ascending and
    columns ascending inside a band.  The input split is not modified."""
    from tike_b200 import cluster
    rng = np.random.default_rng(2)
    scan = np.stack([rng.uniform(1, 900, 4000), rng.uniform(1, 1200, 4000)], 1).astype(np.float32)
    order, batches, start = cluster.by_scan_stripes_contiguous(
        scan=scan, num_workers=3, batch_method='wobbly_center', num_batch=4)
    before = [o.copy() for o in order]
    new = cluster.band_sort_batches(scan, order, batches)
    f = np.floor(scan).astype(np.int64)
    for g in range(3):
        assert np.array_equal(order[g], before[g])
        assert sorted(new[g].tolist()) == sorted(before[g].tolist())
        for b in batches[g]:
            assert np.array_equal(np.sort(new[g][b]), np.sort(before[g][b]))
            band = f[new[g][b], 0] // cluster.BAND_ROWS
            assert np.all(np.diff(band) >= 0)
            for v in np.unique(band):
                assert np.all(np.diff(f[new[g][b], 1][band == v]) >= 0)
    assert sorted(np.concatenate(new).tolist()) == list(range(len(scan)))


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    """Every export validates its arguments before the first CUDA call and
    reports through the return code + tb_last_error (thread-local), never by
    throwing across the ABI; the Python wrapper turns code -1 into ValueError
    (the reference raises ValueError / asserts on bad shapes)."""
    import ctypes
    from tike_b200 import _lib
    h = _lib.lib()
    null = None
    rc = h.tb_patch_fwd(null, null, null, 1, 8, 8, 1, 1, 4, 4, null)
    assert rc == -1 and b'tb_patch_fwd' in h.tb_last_error()
    rc = h.tb_precond_psi(null, 1, 8, null, null, 4, null, 16, 16, null, null)
    assert rc == -1 and b'tb_precond_psi' in h.tb_last_error()
    rc = h.tb_precond_probe(null, 16, 16, null, null, 4, 8, null, null)
    assert rc == -1 and b'tb_precond_probe' in h.tb_last_error()
    rc = h.tb_rpie_update_psi(null, null, null, 10, 0.5, null, null)
    assert rc == -1 and b'tb_rpie_update_psi' in h.tb_last_error()
    count = ctypes.c_int64(-7)
    rc = h.tb_affine_inliers(null, null, null, null, 3, null, 0.0, 0.0, 1.0, null,
                             ctypes.byref(count))
    assert rc == -1 and count.value == -7
    with pytest.raises(ValueError):
        _lib.check(rc, 'affine inliers')
    # an empty problem is fine and does no work
    rc = h.tb_affine_inliers(null, null, null, null, 0,
                             np.zeros(4).ctypes.data, 0.0, 0.0, 1.0, null, ctypes.byref(count))
    assert rc == 0 and count.value == 0


def test_draw_sequence_keeps_the_generator_and_predicts_the_next_epoch():
    """solvers/_common.draw_sequence: the batch orders are exactly what the
    reference draws (rpie.py:95-98: one permutation per epoch from
    tike.random.randomizer_np); the prediction of the next epoch's order is
    obtained by peeking, i.e. it equals the next draw and consumes nothing."""
    import tike_b200.random
    from tike_b200.ptycho.solvers._common import draw_sequence
    tike_b200.random.randomizer_np = np.random.default_rng(11)
    got = [draw_sequence(7, False) for _ in range(4)]
    ref = np.random.default_rng(11)
    want = [[int(x) for x in ref.permutation(7)] for _ in range(5)]
    for e, (seq, nxt) in enumerate(got):
        assert seq == want[e]
        assert nxt == want[e + 1]
    assert draw_sequence(3, True) == ([0, 1, 2], [0, 1, 2])


def test_batch_stager_plan_and_next_epoch_prefetch(monkeypatch):
    """solvers/_common.BatchStager on host data with a stub in place of the CUDA
    upload ring: pieces cover every batch exactly once in sequence order, never
    straddle a batch's cut, at most depth pieces are requested ahead, and
    prefetch_next starts the first pieces of the predicted next epoch."""
    from tike_b200.ptycho.solvers import _common

    class StubRing:
        def __init__(self):
            self.issued, self.taken, self.pending, self.rows = [], [], set(), 0
            self.buffers = [np.zeros((0, 4, 4), np.float32)]

        def issue(self, lo, hi):
            if (lo, hi) not in self.pending:  # like _HostRing: only uploads in flight are known
                self.pending.add((lo, hi))
                self.issued.append((lo, hi))

        def take(self, lo, hi):
            self.issue(lo, hi)
            self.pending.discard((lo, hi))
            self.taken.append((lo, hi))
            return 0, np.zeros((hi - lo, 4, 4), np.float32)

        def release(self, slot):
            pass

    ring = StubRing()
    monkeypatch.setattr(_common, '_ring_for', lambda *a, **k: ring)
    monkeypatch.setenv('TB_STAGE_CHUNK', '7')
    data = np.zeros((60, 4, 4), np.float32)
    batches = [np.arange(0, 25), np.arange(25, 40), np.arange(40, 60)]
    cuts = [10, 25, 47]
    st = _common.BatchStager(data, batches, [2, 0, 1], 'cpu', cuts=cuts)
    assert ring.issued == [(40, 47), (47, 54)]          # depth = 2 pieces ahead
    got = []
    for k in range(3):
        for lo, hi, chunk in st.chunks(k):
            assert chunk.shape[0] == hi - lo
            assert len(ring.pending) <= 2
            got.append((lo, hi))
    assert got == [(40, 47), (47, 54), (54, 60), (0, 7), (7, 10), (10, 17), (17, 24), (24, 25),
                   (25, 32), (32, 39), (39, 40)]
    assert ring.taken == got
    st.prefetch_next([1, 2, 0])
    assert ring.issued[-2:] == [(25, 32), (32, 39)]
    st.prefetch_next(None)                                # no prediction: nothing happens
    assert ring.issued[-2:] == [(25, 32), (32, 39)]
