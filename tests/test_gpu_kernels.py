"""GPU parity tests of the CUDA kernels, called through the C-ABI.

Each test compares libtikeb200 with the CPU oracle (oracle/ptycho_np.py) or
with the committed golden vectors produced by the reference itself
(tests/golden/*.npz).  Tolerances: per-batch intensities and gradients 1e-4
relative L2 (BASELINE.json north_star); patch KATs atol 1e-6 like the
reference's own tests.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope='module')
def K():
    from tike_b200 import kernels
    return kernels


@pytest.fixture(scope='module')
def onp():
    from oracle import ptycho_np
    return ptycho_np


# ------------------------------------------------------------------ patch --
def test_patch_fwd_known_answers(K):
    size, win = 256, 8
    rng = np.random.default_rng(0)
    fov = (rng.random((size, size)) - 0.5 + 1j * (rng.random((size, size)) - 0.5)).astype(np.complex64)
    sub = 0.12346789
    c = size // 2 - win // 2
    positions = np.array([[0, 0], [0, size - win], [size - win, 0],
                          [size - win, size - win], [c, c], [sub, 3]], dtype=np.float32)
    truth = np.stack([
        fov[:win, :win], fov[:win, -win:], fov[-win:, :win], fov[-win:, -win:],
        fov[c:c + win, c:c + win],
        (1.0 - sub) * fov[0:win, 3:3 + win] + sub * fov[1:1 + win, 3:3 + win]])
    patches = torch.zeros((6, win, win), dtype=torch.complex64, device='cuda')
    K.patch_fwd(dev(fov), dev(positions), patches, win)
    np.testing.assert_allclose(host(patches), truth, atol=1e-6)


def test_patch_adj_known_answers(K):
    size, win = 8, 2
    positions = np.array([[0, 0], [0, size - win], [size - win, 0],
                          [size - win, size - win], [3, 3], [0.123, 3],
                          [3, 0.123], [5.5, 3.5]], dtype=np.float32)
    fov = np.zeros((size, size), dtype=np.complex64)
    for (y, x, w) in [(0, 0, 1), (0, size - win, 1), (size - win, 0, 1),
                      (size - win, size - win, 1), (3, 3, 1), (0, 3, 1 - 0.123),
                      (1, 3, 0.123), (3, 0, 1 - 0.123), (3, 1, 0.123),
                      (5, 3, .25), (6, 3, .25), (5, 4, .25), (6, 4, .25)]:
        fov[y:y + win, x:x + win] += w
    images = torch.zeros((size, size), dtype=torch.complex64, device='cuda')
    K.patch_adj(images, dev(positions),
                torch.ones((8, win, win), dtype=torch.complex64, device='cuda'), win)
    np.testing.assert_allclose(host(images), fov, atol=1e-6)


@pytest.mark.parametrize('nrepeat,pad', [(1, 0), (3, 0), (2, 4)])
def test_patch_random_vs_oracle(K, onp, nrepeat, pad):
    rng = np.random.default_rng(3)
    H, W, N, B = 70, 90, 12, 33
    img = (rng.standard_normal((H, W)) + 1j * rng.standard_normal((H, W))).astype(np.complex64)
    pos = np.stack([rng.uniform(0, H - N - 1, B), rng.uniform(0, W - N - 1, B)], 1).astype(np.float32)
    ref = np.zeros((B * nrepeat, N + 2 * pad, N + 2 * pad), np.complex64)
    onp.patch_fwd(img, pos, N, nrepeat=nrepeat, patches=ref)
    out = torch.zeros(ref.shape, dtype=torch.complex64, device='cuda')
    K.patch_fwd(dev(img), dev(pos), out, N, nrepeat)
    assert rel_err(host(out), ref) < 1e-6
    # adjoint, including the broadcast of a single patch (K = 1)
    pat = (rng.standard_normal(ref.shape) + 1j * rng.standard_normal(ref.shape)).astype(np.complex64)
    ref_img = onp.patch_adj(pos, pat, np.zeros((H, W), np.complex64), N, nrepeat=nrepeat)
    out_img = torch.zeros((H, W), dtype=torch.complex64, device='cuda')
    K.patch_adj(out_img, dev(pos), dev(pat), N, nrepeat)
    assert rel_err(host(out_img), ref_img) < 1e-5
    if nrepeat == 1 and pad == 0:
        ref_img = onp.patch_adj(pos, pat[:1], np.zeros((H, W), np.complex64), N)
        out_img.zero_()
        K.patch_adj(out_img, dev(pos), dev(pat[:1]), N, 1)
        assert rel_err(host(out_img), ref_img) < 1e-5


def test_patch_adjoint_identity(K):
    """<F m, d> == <m, F* d> (tests/operators/util.py:42-54, rtol 1e-3)."""
    rng = np.random.default_rng(4)
    H, W, N, B = 64, 64, 16, 40
    m = dev((rng.standard_normal((H, W)) + 1j * rng.standard_normal((H, W))).astype(np.complex64))
    d = dev((rng.standard_normal((B, N, N)) + 1j * rng.standard_normal((B, N, N))).astype(np.complex64))
    pos = dev(np.stack([rng.uniform(1, H - N - 2, B), rng.uniform(1, W - N - 2, B)], 1).astype(np.float32))
    Fm = torch.zeros_like(d)
    K.patch_fwd(m, pos, Fm, N)
    Fd = torch.zeros_like(m)
    K.patch_adj(Fd, pos, d, N)
    a = torch.sum(Fm * d.conj()).item()
    b = torch.sum(m * Fd.conj()).item()
    assert abs(a - b) / abs(a) < 1e-3


# -------------------------------------------------------------------- fft --
@pytest.mark.parametrize('n', [16, 32, 64, 128, 256, 512, 1024, 8, 12, 24, 96, 100, 150, 384])
def test_fft2_vs_torch(K, n):
    rng = np.random.default_rng(n)
    batch = 5 if n <= 256 else 2
    x = (rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))).astype(np.complex64)
    xd = dev(x)
    ref = torch.fft.fft2(xd, norm='ortho')
    y = xd.clone()
    K.fft2(y, inverse=False, scale=1.0 / n)
    assert rel_err(host(y), host(ref)) < 2e-6
    K.fft2(y, inverse=True, scale=1.0 / n)
    assert rel_err(host(y), x) < 3e-6


# ---------------------------------------------------------------- forward --
def test_simulate_golden(K):
    """tests/ptycho/test_ptycho.py:191-203 (ptycho_setup.pickle.lzma)."""
    g = load_golden('ptycho_setup')
    psi, probe, scan = dev(g['psi'][0]), dev(g['probe'][0, 0]), dev(g['scan'])
    b = K.make_batch(psi, scan, probe, 32)
    inten = torch.empty((len(g['scan']), 32, 32), dtype=torch.float32, device='cuda')
    K.ptycho_fwd(b, None, inten)
    np.testing.assert_allclose(np.sqrt(host(inten)), np.sqrt(g['data']), atol=1e-6)


@pytest.mark.parametrize('det,N,M', [(16, 16, 3), (32, 16, 2), (64, 64, 2), (128, 128, 2), (128, 96, 1), (256, 256, 1)])
def test_farplane_vs_oracle(K, onp, det, N, M):
    from tike_b200 import synthetic
    H, W = N + 40, N + 52
    psi, probe, scan = synthetic.make_problem(9, N, M, H, W, seed=det + M)
    far_ref = onp.farplane(psi, scan, probe, det)
    b = K.make_batch(dev(psi[0]), dev(scan), dev(probe[0, 0]), det)
    far = torch.empty((9, M, det, det), dtype=torch.complex64, device='cuda')
    inten = torch.empty((9, det, det), dtype=torch.float32, device='cuda')
    K.ptycho_fwd(b, far, inten)
    assert rel_err(host(far), far_ref[:, 0]) < 1e-5
    assert rel_err(host(inten), onp.intensity(far_ref)) < 1e-5


# ------------------------------------------------------------------- rPIE --
def _rpie_gpu(K, g):
    det = int(g['det'])
    psi, probe = dev(g['psi']), dev(g['probe'])
    scan, data = dev(g['scan']), dev(g['data'])
    mask = g['mask']
    ew = dev(g['eigen_weights']) if g['eigen_weights'].size else None
    b = K.make_batch(psi[0], scan, probe[0, 0], det, eigen_weights=ew)
    B = scan.shape[0]
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    psi_num = torch.zeros_like(psi)
    probe_num = torch.empty_like(probe[0, 0])
    step = torch.empty(B, dtype=torch.float32, device='cuda') if ew is not None else None
    K.rpie_batch(b, data, None if mask.all() else dev(mask.astype(np.uint8)),
                 int(mask.sum()), noise_model=str(g['noise_model']),
                 step_mode=str(g['usemodes']),
                 unmeasured_scaling=float(g['scaling']), psi_numerator=psi_num[0],
                 probe_numerator=probe_num, costs=costs, eigen_weight_step=step)
    torch.cuda.synchronize()
    return psi, probe, scan, costs, psi_num, probe_num, step, ew


@pytest.mark.parametrize('tag', ['rpie_batch_a', 'rpie_batch_pad', 'rpie_batch_poisson',
                                 'rpie_batch_poisson_dom', 'rpie_batch_eigen'])
def test_rpie_batch_golden(K, tag):
    g = load_golden(tag)
    psi, probe, scan, costs, psi_num, probe_num, step, ew = _rpie_gpu(K, g)
    assert rel_err(host(costs), g['costs']) < TOL
    assert rel_err(host(psi_num), g['psi_num']) < TOL
    assert rel_err(host(probe_num), g['probe_num'][0, 0, 0]) < TOL
    if ew is not None:
        new = g['eigen_weights'].copy()
        new[:, 0, 0] += host(step)
        assert rel_err(new, g['eigen_weights_out']) < TOL
    # preconditioners and the update
    pp = torch.empty_like(psi)
    K.precond_psi(probe[0, 0], scan, pp[0])
    qp = torch.empty((1, *probe.shape[-2:]), dtype=torch.complex64, device='cuda')
    K.precond_probe(psi[0], scan, qp[0])
    assert rel_err(host(pp), g['psi_precond']) < TOL
    assert rel_err(host(qp), g['probe_precond']) < TOL
    psi_new, probe_new = psi.clone(), probe.clone()
    K.rpie_update_psi(psi_new, dev(g['psi_num']), dev(g['psi_precond']), float(g['alpha']))
    K.rpie_update_probe(probe_new, dev(g['probe_num'][0]), dev(g['probe_precond']), float(g['alpha']))
    assert rel_err(host(psi_new), g['psi_new']) < 1e-5
    assert rel_err(host(probe_new), g['probe_new']) < 1e-5


@pytest.mark.parametrize('det,N,M,B', [(64, 64, 3, 20), (128, 128, 2, 9), (128, 100, 8, 5),
                                       (256, 256, 2, 3), (256, 200, 1, 2), (512, 512, 1, 2),
                                       (128, 128, 16, 3), (64, 40, 3, 9), (32, 20, 2, 11),
                                       (128, 101, 2, 4), (1024, 1024, 1, 1),
                                       (2048, 2048, 1, 1),
                                       # the exact headline tile of bench.py (BASELINE
                                       # configs[1]): rpie_p3_kernel (three-pass), M = 8;
                                       # 160 positions > 148 SMs, so some persistent CTAs
                                       # take a second position
                                       (128, 128, 8, 16), (128, 128, 8, 160),
                                       # large-detector pipeline at the mode counts of
                                       # BASELINE configs 3 and 5
                                       (256, 256, 4, 3), (512, 512, 2, 2)])
def test_rpie_batch_vs_oracle_large(K, onp, det, N, M, B):
    """Same check at the fused kernel's production tile sizes."""
    from tike_b200 import synthetic
    H, W = N + 60, N + 70
    psi_t, probe, scan = synthetic.make_problem(B, N, M, H, W, seed=det)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(1)
    psi = (psi_t * (1 + 0.1 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch(data, scan, psi, probe, mask)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model='gaussian', usemodes='all_modes',
             scaling=1.0)
    _, _, _, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[0, 0, 0]) < TOL


def test_large_k2_tma_ring_equals_register_lookahead(K, monkeypatch):
    """K2 of the 256 x 256 pipeline (csrc/large_k2r.cu) with its input tiles
    through the TMA ring (cp.async.bulk + mbarrier, TB_LARGE_K2R_TMA=1) and through
    registers: same arithmetic, so the results agree to float-atomic summation
    order; more row blocks than persistent CTAs, so the ring wraps."""
    from tike_b200 import synthetic
    det = N = 256
    M, B = 3, 40
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 60, N + 70, seed=5)
    g = torch.Generator(device='cuda').manual_seed(0)
    data = torch.rand((B, det, det), device='cuda', generator=g) * 50
    out = []
    for ring in ('0', '1'):
        monkeypatch.setenv('TB_LARGE_K2R_TMA', ring)
        psi_d, probe_d, scan_d = dev(psi_t), dev(probe), dev(scan)
        b = K.make_batch(psi_d[0], scan_d, probe_d[0, 0], det)
        costs = torch.empty(B, dtype=torch.float32, device='cuda')
        psi_num = torch.zeros_like(psi_d[0])
        probe_num = torch.empty_like(probe_d[0, 0])
        K.rpie_batch(b, data, None, det * det, noise_model='gaussian',
                     psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
        torch.cuda.synchronize()
        out.append((host(costs), host(probe_num), host(psi_num)))
    assert rel_err(out[0][0], out[1][0]) < 1e-6  # sums of float atomics in any order
    assert rel_err(out[0][1], out[1][1]) < 1e-6
    assert rel_err(out[0][2], out[1][2]) < 1e-6


@pytest.mark.parametrize('M,W,B', [(1, 198, 7), (3, 197, 5), (5, 256, 151)])
def test_rpie_three_pass_kernel_corner_cases(K, onp, M, W, B):
    """rpie_p3_kernel (csrc/rpie_p3.cu) beyond the headline shape: a single mode
    (nothing is spilled, the only far field is parked in Tensor Memory), an odd
    object width (no 16-byte aligned window: the bilinear patch comes from
    global loads instead of the TMA window), an odd mode count, and more
    positions than persistent CTAs; costs and both numerators against the
    oracle."""
    from tike_b200 import synthetic
    det = N = 128
    H = N + 60
    psi_t, probe, scan = synthetic.make_problem(B, N, M, H, W, seed=M + W)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(2)
    psi = (psi_t * (1 + 0.1 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch(data, scan, psi, probe, mask)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model='gaussian', usemodes='all_modes',
             scaling=1.0)
    _, _, _, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[0, 0, 0]) < TOL


@pytest.mark.parametrize('det,M,E,Me', [(32, 2, 1, 1), (64, 3, 2, 2), (128, 2, 1, 2), (128, 3, 0, 0)])
def test_rpie_varying_probe_vs_oracle(K, onp, det, M, E, Me):
    """Per-position varying probe (weights + eigen probes) and the eigen-weight
    step through the stage-fused kernel (probe width == detector width)."""
    from tike_b200 import synthetic
    N, B = det, 7
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 60, N + 70, seed=det + M)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(3)
    psi = (psi_t * (1 + 0.1 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    weights = (1 + 0.2 * rng.standard_normal((B, E + 1, M))).astype(np.float32)
    eigen = None
    if E:
        eigen = (0.1 * np.abs(probe).max() * (rng.standard_normal((1, E, Me, N, N)) +
                 1j * rng.standard_normal((1, E, Me, N, N)))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    c_ref, pn_ref, qn_ref, ew_ref = onp.rpie_batch(data, scan, psi, probe, mask,
                                                   eigen_probe=eigen, eigen_weights=weights)
    psi_d, probe_d, scan_d, data_d = dev(psi), dev(probe), dev(scan), dev(data)
    b = K.make_batch(psi_d[0], scan_d, probe_d[0, 0], det,
                     eigen_probe=dev(eigen[0]) if E else None, eigen_weights=dev(weights))
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    psi_num = torch.zeros_like(psi_d)
    probe_num = torch.empty_like(probe_d[0, 0])
    step = torch.empty(B, dtype=torch.float32, device='cuda')
    K.rpie_batch(b, data_d, None, det * det, noise_model='gaussian', psi_numerator=psi_num[0],
                 probe_numerator=probe_num, costs=costs, eigen_weight_step=step)
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[0, 0, 0]) < TOL
    assert rel_err(weights[:, 0, 0] + host(step), ew_ref[:, 0, 0]) < TOL


@pytest.mark.parametrize('det,N,noise,usemodes', [(256, 256, 'poisson', 'all_modes'),
                                                   (256, 192, 'poisson', 'dominant_mode'),
                                                   (64, 64, 'poisson', 'all_modes'),
                                                   (32, 32, 'poisson', 'dominant_mode'),
                                                   (128, 128, 'poisson', 'all_modes'),
                                                   (128, 128, 'poisson', 'dominant_mode')])
def test_rpie_poisson_vs_oracle(K, onp, det, N, noise, usemodes):
    """Poisson step lengths at production tile sizes, incl. the two-pass path."""
    from tike_b200 import synthetic
    M, B = 2, 3
    H, W = N + 60, N + 70
    psi_t, probe, scan = synthetic.make_problem(B, N, M, H, W, seed=det + 1)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(2)
    psi = (psi_t * (1 + 0.05 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    mask[5:9, :] = False
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch(data, scan, psi, probe, mask, noise_model=noise,
                                              usemodes=usemodes, unmeasured_scaling=0.8)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model=noise, usemodes=usemodes, scaling=0.8)
    _, _, _, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    # the Poisson cost mean(I - d log I) cancels heavily in float32: 1e-3 (the
    # north-star bar for costs); gradients keep the 1e-4 bar
    assert rel_err(host(costs), c_ref) < 1e-3
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[0, 0, 0]) < TOL


def test_lstsq_large_detector_vs_oracle(K, onp):
    """lstsq phase 1 + 2 on a 256^2 detector (two-pass FFT path)."""
    from tike_b200 import synthetic
    from tike_b200.ptycho.position import gaussian_gradient_taps
    det = N = 256
    M, B = 2, 3
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 60, N + 70, seed=9)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(4)
    psi = (psi_t * (1 + 0.1 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    pre = onp.psi_preconditioner(psi, probe, scan)
    r = onp.lstsq_batch(data, scan, psi, probe, mask, pre, 2, recover_positions=True)
    psi_d, probe_d, scan_d, data_d = dev(psi), dev(probe), dev(scan), dev(data)
    b = K.make_batch(psi_d[0], scan_d, probe_d[0, 0], det)
    chi = torch.empty((B, M, N, N), dtype=torch.complex64, device='cuda')
    obj = torch.zeros_like(psi_d)
    psum = torch.empty((M, N, N), dtype=torch.complex64, device='cuda')
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    pnum = torch.zeros((B, 2), dtype=torch.float32, device='cuda')
    pden = torch.zeros((B, 2), dtype=torch.float32, device='cuda')
    K.lstsq_phase1(b, data_d, None, det * det, noise_model='gaussian', chi=chi,
                   object_upd_sum=obj[0], probe_upd_sum=psum, costs=costs,
                   position_num=pnum, position_den=pden, taps=gaussian_gradient_taps())
    assert rel_err(host(chi), r['chi'][:, 0]) < TOL
    assert rel_err(host(obj), r['object_upd_sum']) < TOL
    assert rel_err(host(psum) / 2, r['m_probe_update'][0, 0]) < TOL
    assert rel_err(host(costs), r['costs']) < TOL
    assert rel_err(host(pnum), r['pos_num']) < 1e-3
    pp = torch.empty_like(psi_d)
    K.precond_psi(probe_d[0, 0], scan_d, pp[0])
    assert rel_err(host(pp), pre) < TOL
    qp = torch.empty((1, N, N), dtype=torch.complex64, device='cuda')
    K.precond_probe(psi_d[0], scan_d, qp[0])
    assert rel_err(host(qp), onp.probe_preconditioner(psi, probe, scan)) < TOL


def test_rpie_uint16_data(K, onp):
    from tike_b200 import synthetic
    det = N = 32
    psi, probe, scan = synthetic.make_problem(12, N, 2, 80, 90, seed=5)
    data = np.round(onp.simulate(det, probe, scan, psi) * 20).astype(np.uint16)
    mask = np.ones((det, det), bool)
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch(data.astype(np.float32), scan, psi, probe, mask)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model='gaussian', usemodes='all_modes', scaling=1.0)
    _, _, _, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL


# ------------------------------------------------------------------ lstsq --
@pytest.mark.parametrize('tag', ['lstsq_batch_a', 'lstsq_batch_pad', 'lstsq_batch_poisson'])
def test_lstsq_batch_golden(K, tag):
    import scipy.ndimage
    g = load_golden(tag)
    det = int(g['det'])
    psi, probe = dev(g['psi']), dev(g['probe'])
    scan, data = dev(g['scan']), dev(g['data'])
    B, M, N = scan.shape[0], probe.shape[-3], probe.shape[-1]
    b = K.make_batch(psi[0], scan, probe[0, 0], det)
    chi = torch.empty((B, M, N, N), dtype=torch.complex64, device='cuda')
    obj = torch.zeros_like(psi)
    psum = torch.empty((M, N, N), dtype=torch.complex64, device='cuda')
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    pnum = torch.zeros((B, 2), dtype=torch.float32, device='cuda')
    pden = torch.zeros((B, 2), dtype=torch.float32, device='cuda')
    imp = np.zeros(5, np.float32)
    imp[2] = 1
    taps = scipy.ndimage.gaussian_filter1d(imp, sigma=0.333, order=1, mode='constant', truncate=6.0)[::-1]
    K.lstsq_phase1(b, data, None, det * det, noise_model=str(g['noise_model']), chi=chi,
                   object_upd_sum=obj[0], probe_upd_sum=psum, costs=costs,
                   position_num=pnum, position_den=pden, taps=taps)
    assert rel_err(host(chi), g['chi'][:, 0]) < TOL
    assert rel_err(host(obj), g['obj_sum']) < TOL
    assert rel_err(host(psum) / int(g['num_batch']), g['m_probe_update'][0, 0]) < TOL
    assert rel_err(host(costs), g['costs']) < TOL
    assert rel_err(host(pnum), g['pos_num']) < 1e-3
    assert rel_err(host(pden), g['pos_den']) < 1e-3
    precond = torch.empty_like(psi)
    K.lstsq_precondition_object(precond, dev(g['obj_sum']), dev(g['psi_precond']))
    assert rel_err(host(precond), g['precond']) < 1e-5
    out = torch.empty((B, 6), dtype=torch.float32, device='cuda')
    mpu = dev(g['m_probe_update'][0, 0, 0])
    K.lstsq_phase2(b, dev(g['chi'][:, 0]), dev(g['precond'][0]), mpu, 0, 1e-9 / (N * N), out)
    o = host(out).astype(np.float64)
    A1, A4, b1, b2 = o[:, 0], o[:, 1], o[:, 2], o[:, 3]
    A2 = o[:, 4] + 1j * o[:, 5]
    A1 = A1 + 0.5 * A1.mean()
    A4 = A4 + 0.5 * A4.mean()
    detm = A1 * A4 - A2 * A2.conj()
    x1 = -np.conj(A2 * b2 - A4 * b1) / detm
    x2 = np.conj(A1 * b2 - A2.conj() * b1) / detm
    beta_o = np.mean(0.9 * np.maximum(0, x1.real))
    beta_p = np.mean(0.9 * np.maximum(0, x2.real))
    assert abs(beta_o - float(g['beta_object'].ravel()[0])) / abs(float(g['beta_object'].ravel()[0])) < 1e-3
    assert abs(beta_p - float(g['beta_probe'].ravel()[0])) / abs(float(g['beta_probe'].ravel()[0])) < 1e-3


def test_lstsq_eigen_probe_batch_golden(K):
    """lstsq_grad with a varying probe, one batch: phase 1, the step-length
    solve on the probe snapshot taken BEFORE the eigen update, and the fused
    eigen-probe / eigen-weight update (csrc/eigen.cu) against the reference's own
    _get_nearplane_gradients / _precondition_nearplane_gradients /
    _update_nearplane (lstsq.py:297-364, probe.py:362-476)."""
    from tike_b200.ptycho.solvers import lstsq as L
    g = load_golden('lstsq_batch_eigen')
    det, nb = int(g['det']), int(g['num_batch'])
    psi, probe = dev(g['psi']), dev(g['probe'])
    scan, data = dev(g['scan']), dev(g['data'])
    ep, ew = dev(g['eigen_probe']), dev(g['eigen_weights'])
    B, M, N = scan.shape[0], probe.shape[-3], probe.shape[-1]
    b = K.make_batch(psi[0], scan, probe[0, 0], det, eigen_probe=ep[0], eigen_weights=ew)
    chi = torch.empty((B, 1, M, N, N), dtype=torch.complex64, device='cuda')
    obj = torch.zeros_like(psi)
    psum = torch.empty_like(probe)
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    K.lstsq_phase1(b, data, None, det * det, noise_model='gaussian', chi=chi,
                   object_upd_sum=obj[0], probe_upd_sum=psum[0, 0], costs=costs)
    assert rel_err(host(chi), g['chi']) < TOL
    assert rel_err(host(obj), g['obj_sum']) < TOL
    mpu = psum / nb
    assert rel_err(host(mpu), g['m_probe_update']) < TOL
    _, beta_o, beta_p = L._precondition_nearplane_gradients(
        b, chi, obj, mpu, dev(g['psi_precond']), recover_psi=True, recover_probe=True)
    assert abs(float(beta_o) / float(g['beta_object'].ravel()[0]) - 1) < 1e-3
    assert abs(float(beta_p) / float(g['beta_probe'].ravel()[0]) - 1) < 1e-3
    ep_new, ew_new = L._update_nearplane(chi, mpu, probe, psi, scan, ep.clone(), ew.clone(),
                                         0, B, num_batch=nb)
    assert rel_err(host(ep_new), g['eigen_probe_new']) < 2e-4
    assert rel_err(host(ew_new), g['eigen_weights_new']) < 2e-4
    # the update moved something
    assert rel_err(g['eigen_probe_new'], g['eigen_probe']) > 1e-3


def test_library_rejects_bad_arguments(K):
    x = torch.zeros((2, 1500, 1500), dtype=torch.complex64, device='cuda')
    with pytest.raises(ValueError):
        K.fft2(x)  # widths that are not powers of two are supported up to 1024
    with pytest.raises(TypeError):
        K.fft2(np.zeros((2, 16, 16), np.complex64))  # host arrays are refused


# ------------------------------------------------------------- multislice --
def _ms_gpu(K, psi, probe, scan, data, h, det):
    psi_d, probe_d, scan_d, data_d = dev(psi), dev(probe), dev(scan), dev(data)
    h_d = dev(h)
    D, B = psi.shape[0], scan.shape[0]
    M = probe.shape[-3]
    b = K.multislice_batch(psi_d, scan_d, probe_d[0, 0], det)
    far = torch.empty((B, 1, M, det, det), dtype=torch.complex64, device='cuda')
    inten = torch.empty((B, det, det), dtype=torch.float32, device='cuda')
    K.multislice_fwd(b, D, h_d, far, inten)
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    psi_num = torch.zeros_like(psi_d)
    probe_num = torch.empty((D, M, det, det), dtype=torch.complex64, device='cuda')
    K.multislice_rpie_batch(b, D, h_d, data_d, None, det * det, noise_model='gaussian',
                            psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
    pre = torch.empty_like(psi_d)
    K.multislice_precond_psi(b, D, h_d, pre)
    qre = torch.empty((D, det, det), dtype=torch.complex64, device='cuda')
    for t in range(D):
        K.precond_probe(psi_d[t], scan_d, qre[t])
    torch.cuda.synchronize()
    return far, inten, costs, psi_num, probe_num, pre, qre


def test_multislice_matches_reference(K, onp):
    """Three-slice object: forward model, rPIE numerators and preconditioners
    against the reference golden (rpie.py:374, 441-474)."""
    g = load_golden('rpie_batch_ms')
    det = int(g['det'])
    h = K.fresnel_propagator(det, tuple(g['fov']), float(g['distance']), float(g['wavelength']))
    far, inten, costs, psi_num, probe_num, pre, qre = _ms_gpu(
        K, g['psi'], g['probe'], g['scan'], g['data'], h, det)
    assert rel_err(host(far), g['farplane']) < TOL
    assert rel_err(host(inten), onp.intensity(g['farplane'])) < TOL
    assert rel_err(host(costs), g['costs']) < TOL
    assert rel_err(host(psi_num), g['psi_num']) < TOL
    assert rel_err(host(probe_num), g['probe_num'][:, 0, 0]) < TOL
    assert rel_err(host(pre), g['psi_precond']) < TOL
    assert rel_err(host(qre), g['probe_precond']) < TOL


@pytest.mark.parametrize('det,M,D,B', [(64, 2, 2, 9), (128, 3, 3, 5), (256, 1, 2, 3),
                                       (96, 2, 2, 4),
                                       # per-position fused slice loop (multislice_fused.cu):
                                       # probe = detector width in {32, 64, 128}, D accumulators
                                       # in Tensor Memory; 160 positions > 148 persistent CTAs
                                       (128, 2, 2, 7), (128, 8, 2, 160), (64, 3, 4, 11),
                                       (32, 2, 3, 40)])
def test_multislice_vs_oracle_large(K, onp, det, M, D, B):
    """Same at production tile sizes, against the oracle."""
    from tike_b200 import synthetic
    N = det
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 60, N + 70, seed=det + D)
    rng = np.random.default_rng(7)
    psi = np.stack([(psi_t[0] * (1 + 0.1 * rng.standard_normal(psi_t[0].shape)) *
                     np.exp(0.2j * rng.standard_normal(psi_t[0].shape))).astype(np.complex64)
                    for _ in range(D)])
    fov, dist, lam = (N * 2e-8, N * 2e-8), 3e-6, 1.5e-10
    h = onp.fresnel_propagator(N, fov, dist, lam)
    data = onp.intensity(onp.multislice_farplane(psi, scan, probe, h))
    data = (data * (1 + 0.3 * rng.random(data.shape))).astype(np.float32)
    mask = np.ones((det, det), bool)
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch_multislice(data, scan, psi, probe, mask, h)
    far, inten, costs, psi_num, probe_num, pre, qre = _ms_gpu(
        K, psi, probe, scan, data, K.fresnel_propagator(N, fov, dist, lam), det)
    assert rel_err(host(far), onp.multislice_farplane(psi, scan, probe, h)) < TOL
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[:, 0, 0]) < TOL
    assert rel_err(host(pre), onp.psi_preconditioner_multislice(psi, probe, scan, h)) < TOL
    assert rel_err(host(qre), onp.probe_preconditioner_multislice(psi, probe, scan)) < TOL


def test_multislice_fused_equals_chunked_chain(K, monkeypatch):
    """The per-position fused slice loop and the chunked per-slice chain
    (TB_MULTISLICE_UNFUSED=1) are two implementations of the same batch."""
    from tike_b200 import synthetic
    det = N = 64
    M, D, B = 3, 2, 200
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 90, N + 100, seed=5)
    rng = np.random.default_rng(3)
    psi = np.stack([(psi_t[0] * np.exp(0.2j * rng.standard_normal(psi_t[0].shape))).astype(np.complex64)
                    for _ in range(D)])
    h = K.fresnel_propagator(N, (N * 2e-8, N * 2e-8), 3e-6, 1.5e-10)
    data = (rng.random((B, det, det)) * 50).astype(np.float32)
    out = {}
    for name, flag in (('fused', '0'), ('chunked', '1')):
        monkeypatch.setenv('TB_MULTISLICE_UNFUSED', flag)
        out[name] = _ms_gpu(K, psi, probe, scan, data, h, det)
    # costs, psi_num, probe_num, object preconditioner
    for a, b in zip(out['fused'][2:6], out['chunked'][2:6]):
        assert rel_err(host(a), host(b)) < 2e-5


@pytest.mark.parametrize('det,M', [(32, 2), (64, 3), (128, 2), (256, 2)])
def test_rpie_masked_nan_pixels_fast_paths(K, onp, det, M):
    """Unmeasured detector pixels hold NaN in the reference's tests
    (tests/ptycho/test_ptycho.py:327-334) and are rescaled by
    unmeasured_pixels_scaling - 1: the stage-fused kernel (probe width ==
    detector width) and the fused large-detector pipeline must select on the
    mask, never multiply by data-derived terms."""
    from tike_b200 import synthetic
    N, B = det, 6
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 60, N + 70, seed=det + 3)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(5)
    psi = (psi_t * (1 + 0.1 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    mask[3:7, :] = False
    mask[:, det // 2 - 2:det // 2 + 1] = False
    data = data.copy()
    data[:, ~mask] = np.nan
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch(data, scan, psi, probe, mask,
                                              unmeasured_scaling=0.9)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model='gaussian', usemodes='all_modes',
             scaling=0.9)
    _, _, _, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    assert np.all(np.isfinite(host(psi_num))) and np.all(np.isfinite(host(probe_num)))
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[0, 0, 0]) < TOL


@pytest.mark.parametrize('det', [64, 128, 256])
def test_rpie_cost_only_call(K, onp, det):
    """Without numerators (ObjectOptions=None) the batch call only evaluates the
    cost; same costs as the full call, nothing else written."""
    from tike_b200 import synthetic
    N, M, B = det, 2, 5
    psi_t, probe, scan = synthetic.make_problem(B, N, M, N + 60, N + 70, seed=det + 9)
    data = onp.simulate(det, probe, scan, psi_t)
    psi = (psi_t * 1.1).astype(np.complex64)
    c_ref, _, _, _ = onp.rpie_batch(data, scan, psi, probe, np.ones((det, det), bool))
    psi_d, probe_d, scan_d, data_d = dev(psi), dev(probe), dev(scan), dev(data)
    b = K.make_batch(psi_d[0], scan_d, probe_d[0, 0], det)
    costs = torch.empty(B, dtype=torch.float32, device='cuda')
    K.rpie_batch(b, data_d, None, det * det, noise_model='gaussian', costs=costs)
    assert rel_err(host(costs), c_ref) < TOL


def test_full_size_against_library_composition_and_additivity(K):
    """BASELINE config 2 tile (128^2 detector, 8 modes, 4096^2 object) at a
    size the CPU oracle cannot reach: the fused kernel must agree with the
    reference's op sequence composed from library GPU ops
    (baseline/torch_standin.py, itself checked against the oracle on the CPU),
    and its numerators must be additive over a split of the batch."""
    from baseline import torch_standin as ts
    from tike_b200 import synthetic
    N, M, H, B = 128, 8, 4096, 3000
    dev_ = 'cuda'
    g = torch.Generator(device=dev_).manual_seed(3)
    amp = 0.8 + 0.2 * torch.rand((H, H), device=dev_, generator=g)
    psi = torch.polar(amp, torch.rand((H, H), device=dev_, generator=g) - 0.5).to(torch.complex64)
    probe = torch.as_tensor(synthetic.make_probe(N, M, seed=2)[0, 0], device=dev_)
    scan = torch.as_tensor(synthetic.make_scan(B, H, H, N, seed=1), device=dev_)
    b = K.make_batch(psi, scan, probe, N)
    data = torch.empty((B, N, N), dtype=torch.float32, device=dev_)
    K.ptycho_fwd(b, None, data)
    data *= 1.0 + 0.3 * torch.rand(data.shape, device=dev_, generator=g)

    def fused(lo, hi, psi_num):
        bb = K.make_batch(psi, scan[lo:hi].contiguous(), probe, N)
        costs = torch.empty(hi - lo, dtype=torch.float32, device=dev_)
        probe_num = torch.empty_like(probe)
        K.rpie_batch(bb, data[lo:hi], None, N * N, noise_model='gaussian',
                     psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
        return costs, probe_num

    psi_num = torch.zeros_like(psi)
    costs, probe_num = fused(0, B, psi_num)
    c_ref, pn_ref, qn_ref = ts.rpie_batch(data, scan, psi, probe)
    torch.cuda.synchronize()
    assert rel_err(host(costs), host(c_ref)) < TOL
    assert rel_err(host(psi_num), host(pn_ref)) < TOL
    assert rel_err(host(probe_num), host(qn_ref)) < TOL
    # additivity: two launches accumulate the same object numerator, and the
    # probe numerators of the halves add up to the whole
    psi_num2 = torch.zeros_like(psi)
    c_a, q_a = fused(0, 1234, psi_num2)
    c_b, q_b = fused(1234, B, psi_num2)
    torch.cuda.synchronize()
    assert rel_err(host(psi_num2), host(psi_num)) < 1e-5
    assert rel_err(host(q_a + q_b), host(probe_num)) < 1e-5
    assert rel_err(host(torch.cat([c_a, c_b])), host(costs)) < 1e-6


@pytest.mark.parametrize('det,N,M', [(96, 96, 2), (100, 80, 2), (192, 192, 1), (24, 24, 3)])
def test_rpie_batch_arbitrary_detector_width(K, onp, det, N, M):
    """Detector widths that are not powers of two (cuFFT in the reference takes
    any width): chirp-z transform behind the same C entry points."""
    from tike_b200 import synthetic
    B = 5
    H, W = N + 60, N + 70
    psi_t, probe, scan = synthetic.make_problem(B, N, M, H, W, seed=det)
    data = onp.simulate(det, probe, scan, psi_t)
    rng = np.random.default_rng(1)
    psi = (psi_t * (1 + 0.1 * rng.standard_normal(psi_t.shape))).astype(np.complex64)
    mask = np.ones((det, det), bool)
    c_ref, pn_ref, qn_ref, _ = onp.rpie_batch(data, scan, psi, probe, mask)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model='gaussian', usemodes='all_modes',
             scaling=1.0)
    psi_d, probe_d, scan_d, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    assert rel_err(host(costs), c_ref) < TOL
    assert rel_err(host(psi_num), pn_ref) < TOL
    assert rel_err(host(probe_num), qn_ref[0, 0, 0]) < TOL
    # forward model (simulate) at the same width
    b = K.make_batch(psi_d[0], scan_d, probe_d[0, 0], det)
    far = torch.empty((B, M, det, det), dtype=torch.complex64, device='cuda')
    inten = torch.empty((B, det, det), dtype=torch.float32, device='cuda')
    K.ptycho_fwd(b, far, inten)
    assert rel_err(host(inten), onp.intensity(onp.farplane(psi, scan, probe, det))) < TOL


def test_empty_batch_is_a_no_op(K):
    """Zero positions: every entry point returns without touching its outputs."""
    N = 32
    psi = torch.ones((60, 64), dtype=torch.complex64, device='cuda')
    probe = torch.ones((2, N, N), dtype=torch.complex64, device='cuda')
    scan = torch.zeros((0, 2), dtype=torch.float32, device='cuda')
    data = torch.zeros((0, N, N), dtype=torch.float32, device='cuda')
    b = K.make_batch(psi, scan, probe, N)
    costs = torch.zeros(0, dtype=torch.float32, device='cuda')
    psi_num = torch.full_like(psi, 7.0)
    probe_num = torch.full_like(probe, 7.0)
    K.rpie_batch(b, data, None, N * N, noise_model='gaussian', psi_numerator=psi_num,
                 probe_numerator=probe_num, costs=costs)
    K.ptycho_fwd(b, torch.zeros((0, 2, N, N), dtype=torch.complex64, device='cuda'),
                 torch.zeros((0, N, N), dtype=torch.float32, device='cuda'))
    torch.cuda.synchronize()
    assert torch.all(psi_num == 7.0)
    pre = torch.empty_like(psi)
    K.precond_psi(probe, scan, pre)
    assert torch.all(pre == 0)


@pytest.mark.parametrize('det', [64, 256])
def test_colliding_positions_accumulate(K, onp, det):
    """Every position on the same spot: all object-gradient reductions collide
    on the same (N + 1)^2 pixels and all probe-numerator reductions on the
    same entries; the sums must be B times one position's."""
    from tike_b200 import synthetic
    N, M, B = det, 2, 150
    psi_t, probe, scan1 = synthetic.make_problem(1, N, M, N + 60, N + 70, seed=det + 1)
    scan = np.repeat(scan1, B, axis=0).astype(np.float32)
    data = np.repeat(onp.simulate(det, probe, scan1, psi_t), B, axis=0)
    psi = (psi_t * 1.2).astype(np.complex64)
    mask = np.ones((det, det), bool)
    c1, pn1, qn1, _ = onp.rpie_batch(data[:1], scan1, psi, probe, mask)
    g = dict(det=det, psi=psi, probe=probe, scan=scan, data=data, mask=mask,
             eigen_weights=np.zeros(0), noise_model='gaussian', usemodes='all_modes',
             scaling=1.0)
    _, _, _, costs, psi_num, probe_num, _, _ = _rpie_gpu(K, g)
    assert rel_err(host(costs), np.repeat(c1, B)) < TOL
    assert rel_err(host(psi_num), B * pn1) < TOL
    assert rel_err(host(probe_num), B * qn1[0, 0, 0]) < TOL


# ------------------------------------------------ operator seam (adjoints) --
def _inner(a, b):
    return complex(torch.sum(a.conj() * b).item())


def _rand_c(rng, shape):
    return torch.as_tensor((rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
                           .astype(np.complex64), device='cuda')


@pytest.mark.parametrize('det,N', [(32, 32), (32, 20), (64, 64)])
def test_operator_adjoint_identities(det, N):
    """<fwd(x), y> == <x, adj(y)> for Propagation, Convolution (object and
    probe adjoints) and Ptycho, as tests/operators/util.py:42-54 does in the
    reference (rtol 1e-3 there)."""
    from tike_b200 import operators as ops
    rng = np.random.default_rng(det + N)
    B, M, H, W = 7, 2, N + 30, N + 34
    psi = _rand_c(rng, (H, W))
    probe = _rand_c(rng, (1, M, N, N))
    scan = torch.as_tensor((rng.random((B, 2)) * 20 + 2).astype(np.float32), device='cuda')

    with ops.Propagation(detector_shape=det) as prop:
        x, y = _rand_c(rng, (B, M, det, det)), _rand_c(rng, (B, M, det, det))
        a, b = _inner(prop.fwd(x), y), _inner(x, prop.adj(y))
        assert abs(a - b) <= 1e-4 * abs(a)

    with ops.Convolution(probe_shape=N, detector_shape=det, nz=H, n=W) as conv:
        near = _rand_c(rng, (B, M, det, det))
        fwd = conv.fwd(psi=psi, scan=scan, probe=probe)
        a = _inner(fwd, near)
        b = _inner(psi, conv.adj(nearplane=near, scan=scan, probe=probe))
        assert abs(a - b) <= 1e-3 * abs(a)
        # probe adjoint: sum over positions of adj_probe pairs with the shared probe
        c = _inner(probe[0], conv.adj_probe(nearplane=near, scan=scan, psi=psi).sum(0))
        assert abs(a - c) <= 1e-3 * abs(a)

    with ops.Ptycho(detector_shape=det, probe_shape=N, nz=H, n=W) as op:
        far = _rand_c(rng, (B, 1, M, det, det))
        fwd = op.fwd(psi=psi[None], scan=scan, probe=probe[None])
        a = _inner(fwd, far)
        psi_adj, probe_adj = op.adj(farplane=far, probe=probe[None], scan=scan, psi=psi[None])
        b = _inner(psi[None], psi_adj)
        assert abs(a - b) <= 1e-3 * abs(a)
        # the probe adjoint comes back per position: pair every one with the shared probe
        c = _inner(probe[None].expand_as(probe_adj), probe_adj)
        assert abs(a - c) <= 1e-3 * abs(a)


@pytest.mark.parametrize('pw,depth', [(15, 3), (16, 2)])
def test_multislice_operator_adjoint_identity(pw, depth):
    """tests/operators/test_multislice.py of the reference (odd probe width 15,
    several slices): <fwd(psi), y> == <psi, adj_psi(y)> (homogeneity of degree
    `depth` absorbed by the 1 / nslices of multislice.py:194) and the probe
    adjoint, through the operator composition with the Fresnel step."""
    from tike_b200 import operators as ops
    rng = np.random.default_rng(pw)
    B, M, H, W = 9, 2, 64, 70
    psi = _rand_c(rng, (depth, H, W))
    probe = _rand_c(rng, (1, M, pw, pw))
    scan = torch.as_tensor((rng.random((B, 2)) * 40 + 2).astype(np.float32), device='cuda')
    y = _rand_c(rng, (B, M, pw, pw))
    with ops.Multislice(detector_shape=pw, probe_shape=pw, nz=H, n=W, probe_wavelength=1e-10,
                        probe_FOV_lengths=(1e-5, 1e-5),
                        multislice_propagation_distance=1e-8) as op:
        fwd = op.fwd(probe=probe, scan=scan, psi=psi)
        psi_adj, probe_adj = op.adj(nearplane=y, probe=probe, scan=scan, psi=psi)
    a = _inner(fwd, y)
    b = _inner(psi, psi_adj)
    assert abs(a - b) <= 1e-3 * abs(a)
    c = _inner(probe.expand_as(probe_adj), probe_adj)
    assert abs(a - c) <= 1e-3 * abs(a)


@pytest.mark.parametrize('N,M,P,H,W', [(128, 8, 700, 520, 900), (64, 2, 400, 300, 333),
                                       (100, 3, 150, 256, 300), (16, 1, 300, 90, 200),
                                       (256, 1, 40, 400, 420)])
@pytest.mark.parametrize('how', ['band', 'natural', 'shuffled'])
def test_preconditioners_any_visiting_order(K, onp, N, M, P, H, W, how):
    """csrc/precond.cu: the window kernels (N <= 128, an order given) and the
    direct kernels give the sums of _preconditioner.py:48-167 whatever the
    visiting order: band-sorted (the production path), natural (window kernels
    re-anchoring at nearly every position) or shuffled.  A few positions hang
    over the object edge and take the direct path inside the window kernels."""
    rng = np.random.default_rng(N + P)
    scan = np.stack([rng.uniform(1, H - N - 2, P), rng.uniform(1, W - N - 2, P)], 1)
    scan[:6] = [[-0.5, 3.25], [H - N - 0.5, 7.5], [4.75, -0.25], [9.5, W - N - 0.75],
                [-0.25, -0.75], [H - N - 0.25, W - N - 0.5]]
    scan = scan.astype(np.float32)
    probe = (rng.standard_normal((1, 1, M, N, N)) +
             1j * rng.standard_normal((1, 1, M, N, N))).astype(np.complex64)
    psi = (rng.standard_normal((1, H, W)) + 1j * rng.standard_normal((1, H, W))).astype(np.complex64)
    scan_d, probe_d, psi_d = dev(scan), dev(probe), dev(psi)
    if how == 'band':
        order = K.band_order(scan_d)
        o = host(order)
        assert sorted(o.tolist()) == list(range(P))
        f = np.floor(scan).astype(np.int64)
        key = (f[:, 0] // K.PRECOND_BAND) * (1 << 21) + f[:, 1] + (1 << 20)
        assert np.all(np.diff(key[o]) >= 0)
    elif how == 'natural':
        order = torch.arange(P, dtype=torch.int32, device='cuda')
    else:
        order = dev(rng.permutation(P).astype(np.int32))
    pp = torch.empty_like(psi_d)
    qp = torch.empty((1, N, N), dtype=torch.complex64, device='cuda')
    K.precond_psi(probe_d[0, 0], scan_d, pp[0], order=order)
    K.precond_probe(psi_d[0], scan_d, qp[0], order=order)
    pp0 = torch.empty_like(psi_d)
    qp0 = torch.empty_like(qp)
    K.precond_psi(probe_d[0, 0], scan_d, pp0[0])
    K.precond_probe(psi_d[0], scan_d, qp0[0])
    torch.cuda.synchronize()
    assert rel_err(host(pp), onp.psi_preconditioner(psi, probe, scan)) < 1e-5
    assert rel_err(host(qp), onp.probe_preconditioner(psi, probe, scan)) < 1e-5
    assert rel_err(host(pp), host(pp0)) < 1e-5
    assert rel_err(host(qp), host(qp0)) < 1e-5
    assert float(pp.imag.abs().max()) == 0.0 and float(qp.imag.abs().max()) == 0.0
    with pytest.raises(ValueError):
        K.precond_psi(probe_d[0, 0], scan_d, pp[0], order=order[:-1])


# ---------------------------------------------------------------------------
# object-sized update / constraint kernels (csrc/update.cu) against the array
# expressions of the reference restated with torch fp32 ops
# ---------------------------------------------------------------------------
def _randc_np(shape, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * scale).astype(np.complex64)


def test_update_kernels_given_max_equal_local_max(K):
    """rpie / lstsq object steps with the maximum supplied by the caller equal
    the entry points that search the maximum themselves."""
    n = (3, 70, 90)
    psi, num = dev(_randc_np(n, 0)), dev(_randc_np(n, 1))
    pre = dev((np.abs(_randc_np(n, 2)) ** 2).astype(np.complex64))
    for t in range(n[0]):
        a, b = psi[t].clone(), psi[t].clone()
        K.rpie_update_psi(a, num[t], pre[t], 0.2)
        K.rpie_update_psi(b, num[t], pre[t], 0.2, precond_max=K.max_real(pre[t]))
        assert torch.equal(a, b)
        o1, o2 = torch.empty_like(a), torch.empty_like(a)
        K.lstsq_precondition_object(o1, num[t], pre[t], 0.05)
        K.lstsq_precondition_object(o2, num[t], pre[t], 0.05, precond_max=K.max_real(pre[t]))
        assert torch.equal(o1, o2)
    mx = K.max_real(pre[0])
    assert float(mx) == float(pre[0].real.max())
    # a larger, cross-rank maximum changes the step as the formula says
    big = mx * 3
    c = psi[0].clone()
    K.rpie_update_psi(c, num[0], pre[0], 0.2, precond_max=big)
    ref = psi[0] + num[0] / ((1 - 0.2) * pre[0] + 0.2 * big)
    assert rel_err(host(c), host(ref)) < 1e-6


def test_rpie_adam_update_matches_reference_expressions(K):
    """tb_rpie_update_psi_adam = rpie.py:233-267 + opt.adam (opt.py:165-213)."""
    from tike_b200 import opt
    n = (64, 80)
    alpha, vd, md = 0.3, 0.999, 0.9
    psi0 = dev(_randc_np(n, 3))
    pre = dev((np.abs(_randc_np(n, 4)) ** 2 + 0.1).astype(np.complex64))
    psi = psi0.clone()
    v = torch.zeros(n, device='cuda')
    m = torch.zeros(n, dtype=torch.complex64, device='cuda')
    ref, rv, rm = psi0.clone(), None, None
    for it in range(3):
        g = dev(_randc_np(n, 10 + it, 0.1))
        K.rpie_update_psi_adam(psi, g, pre, v, m, alpha, vd, md)
        deno = (1 - alpha) * pre + alpha * pre.real.max()
        ref = ref + g / deno
        d, rv, rm = opt.adam(g=g, v=rv, m=rm, vdecay=vd, mdecay=md)
        ref = ref + d / deno
    assert rel_err(host(psi), host(ref)) < 1e-5
    assert rel_err(host(v), host(rv)) < 1e-5 and rel_err(host(m), host(rm)) < 1e-5


def test_momentum_update_and_add_quotient(K):
    from tike_b200 import opt
    n = (2, 50, 60)
    psi0, x = dev(_randc_np(n, 5)), dev(_randc_np(n, 6))
    beta = torch.tensor([0.37], device='cuda')
    psi, m = psi0.clone(), torch.zeros_like(psi0)
    ref, rm = psi0.clone(), None
    for _ in range(3):
        K.momentum_update(psi, x, m, 0.9, beta)
        d, _, rm = opt.momentum(g=beta * x, v=None, m=rm, mdecay=0.9)
        ref = ref + d
    assert rel_err(host(psi), host(ref)) < 1e-6
    pre = dev((np.abs(_randc_np(n, 7)) ** 2).astype(np.complex64))
    y = psi0.clone()
    K.add_quotient(y, x, pre, 1e-9)
    assert rel_err(host(y), host(psi0 + x / (pre.real + 1e-9))) < 1e-6
    # probe-style call: one (N, N) preconditioner for all modes
    probe, pnum = dev(_randc_np((4, 16, 16), 8)), dev(_randc_np((4, 16, 16), 9))
    ppre = dev((np.abs(_randc_np((16, 16), 10)) ** 2).astype(np.complex64))
    y = probe.clone()
    K.add_quotient(y, pnum, ppre, 1e-9, period=256)
    assert rel_err(host(y), host(probe + pnum / (ppre.real + 1e-9))) < 1e-6


@pytest.mark.parametrize('positivity,smooth,clip', [(0.3, 0.05, True), (0.0, 0.1, False),
                                                    (1.0, 0.0, True), (0.2, 0.0, False)])
def test_object_constraints_match_reference_expressions(K, positivity, smooth, clip):
    """ptycho.py:811-851 through the fused kernels vs ptycho/object.py."""
    import tike_b200.ptycho as tp
    from tike_b200.ptycho import object as tb_object
    from tike_b200.ptycho.ptycho import _apply_object_constraints
    psi0 = dev(_randc_np((2, 37, 53), 11))
    pre = dev((np.abs(_randc_np((2, 37, 53), 12)) ** 2).astype(np.complex64))
    probe0 = dev(_randc_np((1, 1, 2, 8, 8), 13))
    oopt = tp.ObjectOptions(positivity_constraint=positivity, smoothness_constraint=smooth,
                            clip_magnitude=clip)
    oopt.preconditioner = pre
    alg = tp.RpieOptions(rescale_method='mean_of_abs_object', rescale_period=1)
    import types
    p = types.SimpleNamespace(probe=probe0.clone(), psi=psi0.clone(),
                              algorithm_options=alg, object_options=oopt)
    p = _apply_object_constraints(p)
    ref = psi0.clone()
    if positivity:
        ref = tb_object.positivity_constraint(ref, r=positivity)
    if smooth:
        ref = tb_object.smoothness_constraint(ref, a=smooth)
    if clip:
        ref = tb_object.clip_magnitude(ref, a_max=1.0)
    ref, probe_ref = tb_object.remove_object_ambiguity(ref, probe0, pre)
    assert rel_err(host(p.psi), host(ref)) < 2e-6
    assert rel_err(host(p.probe), host(probe_ref)) < 2e-6
