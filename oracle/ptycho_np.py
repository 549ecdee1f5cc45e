"""CPU oracle: a NumPy restatement of the tike ptychography hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package ``tike_b200`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as a
checker or as the CPU baseline.

Parity status: PINNED.  The functions here are checked against
  * the reference's own known-answer tests
    (tests/operators/test_patch.py:64-206, tests/ptycho/test_ptycho.py:191-203
    with tests/data/ptycho_setup.pickle.lzma), and
  * outputs of the unmodified reference solver code executed in the build
    container through the NumPy-backed CuPy shim (oracle/refshim, driver
    tests/golden/make_golden.py -> tests/golden/*.npz).
See tests/test_oracle.py.

All file:line citations are relative to the reference checkout
(AdvancedPhotonSource/tike, multislice fork), src/tike/...

Everything is float32 / complex64 like the reference (precision.py:4-11).
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.ndimage
import scipy.stats

f32 = np.float32
c64 = np.complex64


# ----------------------------------------------------------------------------
# Patch extraction / scatter   (operators/cupy/convolution.cu:35-165,
#                               operators/cupy/patch.py:79-188)
# ----------------------------------------------------------------------------

def _bilinear_terms(positions, height, width, patch_width):
    """Integer corners, the four float32 weights and validity masks.

    convolution.cu:102-133: sy = floor(scan[...,0]), sx = floor(scan[...,1]),
    weights {(1-fx)(1-fy), fx(1-fy), (1-fx)fy, fx*fy} evaluated in float32.
    A patch pixel is skipped when its *leading* image pixel is outside the
    image (:110, :118).  Trailing neighbours outside the image are dropped
    here (the reference dereferences them with zero weight).
    """
    positions = np.asarray(positions, dtype=f32)
    iy = np.floor(positions[:, 0])
    ix = np.floor(positions[:, 1])
    fy = (positions[:, 0] - iy).astype(f32)
    fx = (positions[:, 1] - ix).astype(f32)
    one = f32(1.0)
    w = np.stack(
        [(one - fx) * (one - fy), fx * (one - fy), (one - fx) * fy, fx * fy],
        axis=0,
    ).astype(f32)  # (4, B)
    iy = iy.astype(np.int64)
    ix = ix.astype(np.int64)
    p = np.arange(patch_width, dtype=np.int64)
    yy = iy[:, None, None] + p[None, :, None]  # (B, N, 1)
    xx = ix[:, None, None] + p[None, None, :]  # (B, 1, N)
    yy, xx = np.broadcast_arrays(yy, xx)
    lead_ok = (yy >= 0) & (yy < height) & (xx >= 0) & (xx < width)
    return yy, xx, w, lead_ok


def patch_fwd(images, positions, patch_width, nrepeat=1, patches=None,
              padded_width=None):
    """Bilinear patch extraction.  images (H, W) -> (B*nrepeat, Np, Np).

    patch.py:79-129 + convolution.cu fwd_patch.  The padding border of the
    output (padded_width > patch_width) is left untouched.
    """
    images = np.asarray(images)
    H, W = images.shape[-2:]
    B = len(positions)
    if padded_width is None:
        padded_width = patch_width if patches is None else patches.shape[-1]
    if patches is None:
        patches = np.zeros((B * nrepeat, padded_width, padded_width),
                           dtype=images.dtype)
    pad = (padded_width - patch_width) // 2
    yy, xx, w, lead_ok = _bilinear_terms(positions, H, W, patch_width)

    def take(dy, dx):
        y = yy + dy
        x = xx + dx
        ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
        v = images[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)]
        return np.where(ok, v, 0).astype(images.dtype)

    wv = w[:, :, None, None]
    val = (take(0, 0) * wv[0] + take(0, 1) * wv[1] + take(1, 0) * wv[2] +
           take(1, 1) * wv[3]).astype(images.dtype)
    view = patches.reshape(B, nrepeat, padded_width, padded_width)
    inner = view[:, :, pad:pad + patch_width, pad:pad + patch_width]
    for r in range(nrepeat):
        inner[:, r] = np.where(lead_ok, val, inner[:, r])
    return patches


def patch_adj(positions, patches, images, patch_width, nrepeat=1):
    """Scatter-add patches into images (in place and returned).

    patch.py:131-188 + convolution.cu adj_patch: patch index for position s,
    repeat r is ``r + (nrepeat * s) % K`` with K = number of patches given
    (K < B*nrepeat broadcasts, e.g. _preconditioner.py:70-74).
    """
    images = np.asarray(images)
    H, W = images.shape[-2:]
    B = len(positions)
    K = patches.shape[-3]
    padded_width = patches.shape[-1]
    pad = (padded_width - patch_width) // 2
    assert (B * nrepeat) % K == 0 and K >= nrepeat
    yy, xx, w, lead_ok = _bilinear_terms(positions, H, W, patch_width)
    inner = patches[:, pad:pad + patch_width, pad:pad + patch_width]
    s = np.arange(B)
    flat = images.reshape(-1)
    for r in range(nrepeat):
        src = inner[r + (nrepeat * s) % K]  # (B, N, N)
        for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            y = yy + dy
            x = xx + dx
            ok = lead_ok & (y >= 0) & (y < H) & (x >= 0) & (x < W)
            contrib = (src * w[k][:, None, None]).astype(images.dtype)
            np.add.at(flat, (y * W + x)[ok], contrib[ok])
    return images


# ----------------------------------------------------------------------------
# Forward model   (operators/cupy/convolution.py:58-101, propagation.py:43-73,
#                  ptycho.py(op):114-204, probe.py:272-303)
# ----------------------------------------------------------------------------

def get_varying_probe(shared_probe, eigen_probe=None, weights=None):
    """probe.py:272-303: w0*P + sum_c w_c*E_c  ->  (B, 1, M, N, N)."""
    if weights is None:
        return shared_probe.copy()
    unique = weights[..., [0], :, None, None] * shared_probe
    if eigen_probe is not None:
        m = eigen_probe.shape[-3]
        for c in range(eigen_probe.shape[-4]):
            unique[..., :m, :, :] += (weights[..., [c + 1], :m, None, None] *
                                      eigen_probe[..., [c], :m, :, :])
    return unique.astype(c64)


def exitwave(psi2d, scan, probe, detector_shape):
    """convolution.py:58-101: zero-padded (B, M, Nd, Nd) probe*patch product.

    probe is (B|1, M, N, N)."""
    M, N = probe.shape[-3], probe.shape[-1]
    B = len(scan)
    pad = (detector_shape - N) // 2
    patches = np.zeros((B * M, detector_shape, detector_shape), dtype=c64)
    patches = patch_fwd(psi2d, scan, N, nrepeat=M, patches=patches)
    patches = patches.reshape(B, M, detector_shape, detector_shape)
    patches[..., pad:pad + N, pad:pad + N] *= probe
    return patches


def fft2(x, norm='ortho'):
    """propagation.py:43-57 (cuFFT C2C, DC at the corner)."""
    return scipy.fft.fft2(x, norm=norm, axes=(-2, -1), workers=-1).astype(c64)


def ifft2(x, norm='ortho'):
    """propagation.py:59-73."""
    return scipy.fft.ifft2(x, norm=norm, axes=(-2, -1), workers=-1).astype(c64)


def farplane(psi, scan, probe, detector_shape, norm='ortho'):
    """ptycho.py(op):114-130 for D = 1: (B, 1, M, Nd, Nd).

    psi (1, H, W); probe (B|1, 1, M, N, N)."""
    assert psi.shape[0] == 1, 'oracle covers single-slice objects'
    ew = exitwave(psi[0], scan, probe[..., 0, :, :, :], detector_shape)
    return fft2(ew, norm)[:, None]


def intensity(far):
    """ptycho.py(op):18-23 / rpie.py:376-379: sum over modes of |farplane|^2."""
    return np.sum(np.square(np.abs(far)), axis=tuple(range(1, far.ndim - 2)),
                  dtype=f32)


def simulate(detector_shape, probe, scan, psi, fly=1, eigen_probe=None,
             eigen_weights=None):
    """ptycho.py:95-179: detector intensities, summed mode by mode."""
    scan = np.asarray(scan, dtype=f32)
    psi = np.asarray(psi, dtype=c64)
    probe = np.asarray(probe, dtype=c64)
    out = 0
    for m in range(probe.shape[-3]):
        p = get_varying_probe(
            probe[..., [m], :, :],
            None if eigen_probe is None else eigen_probe[..., [m], :, :],
            None if eigen_weights is None else eigen_weights[..., [m]],
        )
        far = farplane(psi, scan, p, detector_shape)
        out = out + np.sum(
            np.square(np.abs(far)).reshape(len(scan) // fly, fly,
                                           detector_shape, detector_shape),
            axis=1)
    return out.astype(f32)


# ----------------------------------------------------------------------------
# Objective   (operators/cupy/objective.py:11-124)
# ----------------------------------------------------------------------------

def gaussian_each_pattern(data, inten):
    """objective.py:11-15, 47-66 (mean over the trailing axes given).

    The per-pixel terms are float32 like the reference's; the MEAN is
    accumulated in float64: NumPy's float32 sum over a strided axis (which is
    what the boolean-mask gather produces) is a naive running sum and loses
    ~1e-4 at 256^2 pixels, whereas the reference's device reduction is a tree."""
    d = np.sqrt(inten) - np.sqrt(data)
    return np.mean(d * d, axis=(-2, -1), dtype=np.float64).astype(f32)


def poisson_each_pattern(data, inten):
    """objective.py:72-74, 108-124 (float64 accumulation, see above)."""
    return np.mean(inten - data * np.log(inten + f32(1e-9)), axis=(-2, -1),
                   dtype=np.float64).astype(f32)


def pattern_costs(data, inten, mask, noise_model):
    """rpie.py:380-386: per-pattern cost over measured pixels only."""
    d = data[:, mask][:, None, :].astype(f32)
    i = inten[:, mask][:, None, :]
    if noise_model == 'gaussian':
        return gaussian_each_pattern(d, i)
    return poisson_each_pattern(d, i)


def _poisson_steps_all_modes(xi, abs2, I_e, I_m, mask, step, weight):
    """exitwave.py:122-180."""
    I_e = I_e[:, None, None]
    I_m = I_m[:, None, None]
    xa = xi * abs2
    den_final = np.sum((xi * xa)[..., mask], axis=-1)
    for _ in range(2):
        t = xi * step - 1
        den = abs2 * np.square(t) + I_e - abs2
        num = np.sum((xa * (1 + (I_m * t) / den))[..., mask], axis=-1)
        step = step * (1 - weight) + (num / den_final)[..., None, None] * weight
    return step


def _poisson_steps_dominant(xi, I_e, I_m, mask, step, weight):
    """exitwave.py:183-234."""
    I_e = I_e[:, None, None]
    I_m = I_m[:, None, None]
    sden = np.sum((np.square(xi) * I_e)[..., mask], axis=-1)
    for _ in range(2):
        num = xi * (I_e - I_m / (1 - step * xi))
        r = np.sum(num[..., mask], axis=-1) / sden
        step = (1 - weight) * step + weight * r[..., None, None]
    return step


def farplane_gradient(far, data, mask, noise_model='gaussian',
                      unmeasured_scaling=1.0, step_length_start=0.5,
                      step_length_weight=0.5, usemodes='all_modes',
                      poisson_eps_in_xi=False):
    """rpie.py:376-439 / lstsq.py:444-502: returns (chi_hat, costs).

    chi_hat = -farplane * (1 - sqrt(d) / (sqrt(I) + 1e-9)) on measured pixels
    (objective.py:42-44), farplane * (scaling - 1) elsewhere.
    """
    far = far.copy()
    data = np.asarray(data).astype(f32)
    inten = intensity(far)
    costs = pattern_costs(data, inten, mask, noise_model)
    if noise_model == 'poisson':
        if poisson_eps_in_xi:  # lstsq.py:456
            xi = (1 - data / (inten + f32(1e-9)))[:, None, None]
        else:  # rpie.py:390
            xi = (1 - data / inten)[:, None, None]
        grad = far * xi
        step = np.full((far.shape[0], 1, far.shape[2], 1, 1),
                       f32(step_length_start), dtype=f32)
        if usemodes == 'dominant_mode':
            step = _poisson_steps_dominant(xi, inten, data, mask, step,
                                           step_length_weight)
        else:
            step = _poisson_steps_all_modes(xi, np.square(np.abs(far)), inten,
                                            data, mask, step,
                                            step_length_weight)
        far[..., mask] = (-step * grad)[..., mask]
    else:
        g = far * (1 - np.sqrt(data) /
                   (np.sqrt(inten) + f32(1e-9)))[:, None, None]
        far[..., mask] = -g[..., mask]
    far[..., ~mask] *= f32(unmeasured_scaling - 1.0)
    return far.astype(c64), costs.astype(f32)


# ----------------------------------------------------------------------------
# rPIE   (ptycho/solvers/rpie.py)
# ----------------------------------------------------------------------------

def rpie_batch(data, scan, psi, probe, mask, *, eigen_probe=None,
               eigen_weights=None, noise_model='gaussian',
               unmeasured_scaling=1.0, norm='ortho', recover_psi=True,
               recover_probe=True, psi_numerator=None, chunk=64,
               step_length_start=0.5, step_length_weight=0.5,
               usemodes='all_modes'):
    """One call of rpie._get_nearplane_gradients (rpie.py:315-567) for D = 1.

    data (B, Nd, Nd), scan (B, 2), psi (1, H, W), probe (1, 1, M, N, N),
    eigen_weights (B, E+1, M) or None.  Processes the batch in chunks of 64
    like stream_and_modify2 (communicators/stream.py:285-404).

    Returns costs (B,), psi_numerator (1, H, W), probe_numerator
    (1, 1, 1, M, N, N) [re-zeroed every call, rpie.py:349], eigen_weights.
    """
    B = len(scan)
    M, N = probe.shape[-3], probe.shape[-1]
    Nd = data.shape[-1]
    pad = (Nd - N) // 2
    if psi_numerator is None:
        psi_numerator = np.zeros_like(psi)
    probe_numerator = np.zeros((psi.shape[0], *probe.shape), dtype=c64)
    costs = np.empty(B, dtype=f32)
    if eigen_weights is not None:
        eigen_weights = eigen_weights.copy()
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        unique = get_varying_probe(
            probe, eigen_probe,
            eigen_weights[lo:hi] if eigen_weights is not None else None)
        far = farplane(psi, scan[lo:hi], unique, Nd, norm)
        chi_hat, costs[lo:hi] = farplane_gradient(
            far, data[lo:hi], mask, noise_model, unmeasured_scaling,
            step_length_start, step_length_weight, usemodes)
        diff = ifft2(chi_hat, norm)[..., pad:pad + N, pad:pad + N]
        # diff: (b, 1, M, N, N)
        if recover_psi:
            # rpie.py:450-457  conj(probe) * diff / M, scattered with nrepeat=M
            grad_psi = (np.conj(unique[:, 0][:, None]) * diff / M).reshape(
                (hi - lo) * M, N, N).astype(c64)
            psi_numerator[0] = patch_adj(scan[lo:hi], grad_psi,
                                         psi_numerator[0], N, nrepeat=M)
            # rpie.py:459-469
            patches = patch_fwd(psi[0], scan[lo:hi], N)[:, None, None]
            probe_numerator[0] += np.sum(np.conj(patches) * diff, axis=0,
                                         keepdims=True)
        if recover_probe and eigen_weights is not None:
            # rpie.py:477-506 (mode 0 only)
            patches = patch_fwd(psi[0], scan[lo:hi], N)[:, None, None]
            OP = patches * probe[..., 0:1, :, :]
            num = np.sum(np.real(np.conj(OP) * diff[..., 0:1, :, :]),
                         axis=(-1, -2))
            den = np.sum(np.abs(OP)**2, axis=(-1, -2))
            eigen_weights[lo:hi, 0:1, 0:1] += f32(0.1) * (num / den)
    return costs, psi_numerator, probe_numerator, eigen_weights


def adam(g, v=None, m=None, vdecay=0.999, mdecay=0.9, eps=1e-8):
    """opt.py:165-213 (no bias-correction powers, as in the reference)."""
    v = np.zeros_like(g.real) if v is None else v
    m = np.zeros_like(g) if m is None else m
    m = mdecay * m + (1 - mdecay) * g
    v = vdecay * v + (1 - vdecay) * (g * g.conj()).real
    m_ = m / (1 - mdecay)
    v_ = np.sqrt(v / (1 - vdecay))
    return m_ / (v_ + eps), v, m


def rpie_update(psi, probe, psi_numerator, probe_numerator, psi_precond,
                probe_precond, alpha, recover_psi=True, recover_probe=True):
    """rpie._update (rpie.py:217-312) without adaptive moment.

    psi   += G_O / ((1-a) L_O + a max(L_O))
    probe += G_P[0] / (a max(L_P[0]))          (F4: no (1-a) term)
    """
    if recover_psi:
        deno = ((1 - alpha) * psi_precond +
                alpha * psi_precond.max(axis=(-2, -1), keepdims=True))
        psi = (psi + psi_numerator / deno).astype(c64)
    if recover_probe:
        deno = alpha * probe_precond[0].max(axis=(-2, -1), keepdims=True)
        probe = (probe + probe_numerator[0] / deno).astype(c64)
    return psi, probe


def psi_preconditioner(psi, probe, scan, chunk=64):
    """_preconditioner.py:48-104 for D = 1: scatter of sum_m |P_m|^2 (c64)."""
    out = np.zeros(psi.shape, dtype=c64)
    N = probe.shape[-1]
    amp = np.sum(probe * np.conj(probe), axis=-3)[:, 0]  # (1, N, N)
    for lo in range(0, len(scan), chunk):
        out[0] = patch_adj(scan[lo:lo + chunk], amp, out[0], N)
    return out


def probe_preconditioner(psi, probe, scan, chunk=64):
    """_preconditioner.py:116-167: sum_s |patch_s|^2  -> (D, N, N) c64."""
    N = probe.shape[-1]
    out = np.zeros((psi.shape[0], N, N), dtype=c64)
    for lo in range(0, len(scan), chunk):
        patches = patch_fwd(psi[0], scan[lo:lo + chunk], N)
        out[0] += np.sum(patches * np.conj(patches), axis=0)
    return out


def mnorm(x, axis=None, keepdims=False):
    """linalg.py:12-14."""
    return np.sqrt(np.mean((x * x.conj()).real, axis=axis, keepdims=keepdims))


def rpie_epoch(data, scan, psi, probe, mask, batches, order, *, alpha=0.05,
               eigen_probe=None, eigen_weights=None, compact=False,
               recover_psi=True, recover_probe=True, psi_precond=None,
               probe_precond=None, **kw):
    """One call of solvers.rpie (rpie.py:26-206), preconditioners included.

    ``order`` is the sequence of batch indices (the reference draws it from
    tike.random.randomizer_np.permutation, rpie.py:95-98).
    Returns psi, probe, eigen_weights, epoch_cost.
    """
    if psi_precond is None:
        psi_precond = psi_preconditioner(psi, probe, scan)
    if probe_precond is None:
        probe_precond = probe_preconditioner(psi, probe, scan)
    batch_cost = np.zeros(len(batches), dtype=f32)
    psi_num = None
    probe_num = None
    for n in order:
        b = batches[n]
        lo, hi = b[0], b[-1] + 1
        ew = eigen_weights[lo:hi] if eigen_weights is not None else None
        costs, psi_num, probe_num, ew = rpie_batch(
            data[lo:hi], scan[lo:hi], psi, probe, mask,
            eigen_probe=eigen_probe, eigen_weights=ew,
            recover_psi=recover_psi, recover_probe=recover_probe,
            psi_numerator=psi_num, **kw)
        if eigen_weights is not None:
            eigen_weights[lo:hi] = ew
        batch_cost[n] = np.mean(costs)
        if not compact:
            psi, probe = rpie_update(psi, probe, psi_num, probe_num,
                                     psi_precond, probe_precond, alpha,
                                     recover_psi, recover_probe)
            psi_num = None
            probe_num = None
    if compact:
        psi, probe = rpie_update(psi, probe, psi_num, probe_num, psi_precond,
                                 probe_precond, alpha, recover_psi,
                                 recover_probe)
    if eigen_weights is not None:
        eigen_weights = eigen_weights / mnorm(eigen_weights, axis=-3,
                                              keepdims=True)
    return psi, probe, eigen_weights, float(batch_cost.mean())


# ----------------------------------------------------------------------------
# lstsq_grad   (ptycho/solvers/lstsq.py)
# ----------------------------------------------------------------------------

def precondition_object_update(obj_sum, psi_precond, alpha=0.05):
    """lstsq.py:605-616."""
    return (obj_sum / np.sqrt(
        np.square((1 - alpha) * psi_precond) +
        np.square(alpha * np.amax(psi_precond, axis=(-2, -1), keepdims=True)))
           ).astype(c64)


def gaussian_gradient(x, sigma=0.333):
    """position.py:779-810: first-derivative Gaussian along rows / columns."""
    def g(axis):
        return (scipy.ndimage.gaussian_filter1d(
            -x.real, sigma=sigma, order=1, axis=axis, mode='nearest',
            truncate=6.0) + 1j * scipy.ndimage.gaussian_filter1d(
                -x.imag, sigma=sigma, order=1, axis=axis, mode='nearest',
                truncate=6.0)).astype(c64)
    return g(-2), g(-1)


def lstsq_batch(data, scan, psi, probe, mask, psi_precond, num_batch, *,
                eigen_probe=None, eigen_weights=None, noise_model='gaussian',
                unmeasured_scaling=1.0, norm='ortho', recover_psi=True,
                recover_probe=True, recover_positions=False, chunk=64,
                step_length_start=0.5, step_length_weight=0.5,
                usemodes='all_modes'):
    """lstsq._get_nearplane_gradients + _precondition_nearplane_gradients
    (lstsq.py:367-602, 619-718) for one batch, D = 1.

    Returns a dict with chi (B,1,M,N,N), patches (B,1,1,N,N), unique_probe,
    probe_update (per position), object_upd_sum (1,H,W), m_probe_update
    (1,1,M,N,N) [already / num_batch], costs, object_update_precond,
    beta_object, beta_probe, pos_num, pos_den.
    """
    B = len(scan)
    M, N = probe.shape[-3], probe.shape[-1]
    Nd = data.shape[-1]
    pad = (Nd - N) // 2
    chi = np.empty((B, 1, M, N, N), dtype=c64)
    unique = np.empty((B, 1, M, N, N), dtype=c64)
    patches = np.empty((B, 1, 1, N, N), dtype=c64)
    costs = np.empty(B, dtype=f32)
    obj_sum = np.zeros_like(psi)
    pos_num = np.zeros((B, 2), dtype=f32)
    pos_den = np.zeros((B, 2), dtype=f32)
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        unique[lo:hi] = get_varying_probe(
            probe, eigen_probe,
            eigen_weights[lo:hi] if eigen_weights is not None else None)
        far = farplane(psi, scan[lo:hi], unique[lo:hi], Nd, norm)
        chi_hat, costs[lo:hi] = farplane_gradient(
            far, data[lo:hi], mask, noise_model, unmeasured_scaling,
            step_length_start, step_length_weight, usemodes,
            poisson_eps_in_xi=True)
        chi[lo:hi] = ifft2(chi_hat, norm)[..., pad:pad + N, pad:pad + N]
        if recover_psi:
            proj = (np.conj(unique[lo:hi]) * chi[lo:hi]).reshape(
                (hi - lo) * M, N, N)
            obj_sum[0] = patch_adj(scan[lo:hi], proj, obj_sum[0], N, nrepeat=M)
        patches[lo:hi] = patch_fwd(psi[0], scan[lo:hi], N)[:, None, None]
    out = dict(chi=chi, unique_probe=unique, patches=patches, costs=costs,
               object_upd_sum=obj_sum)
    if recover_probe:
        probe_update = (np.conj(patches) * chi).astype(c64)
        m_probe_update = (np.sum(probe_update, axis=0, keepdims=True) /
                          num_batch).astype(c64)
        out.update(probe_update=probe_update, m_probe_update=m_probe_update)
    if recover_positions:
        gx, gy = gaussian_gradient(patches)
        c = N // 4
        sl = (Ellipsis, slice(c, -c), slice(c, -c))
        for k, g in enumerate((gx, gy)):
            gp = g[sl] * unique[:, :, 0:1][sl]
            pos_num[:, k] = np.sum(np.real(np.conj(gp) * chi[:, :, 0:1][sl]),
                                   axis=(-4, -3, -2, -1))
            pos_den[:, k] = np.sum(np.abs(gp)**2, axis=(-4, -3, -2, -1))
        out.update(pos_num=pos_num, pos_den=pos_den)

    # --- _precondition_nearplane_gradients (lstsq.py:619-718), m = 0 ---
    eps = f32(1e-9) / f32(N * N)
    chi0 = chi[..., 0:1, :, :]
    x1 = x2 = None
    if recover_psi:
        precond = precondition_object_update(obj_sum, psi_precond)
        proj = patch_fwd(precond[0], scan, N)
        dOP = proj[:, None, None] * unique[..., 0:1, :, :]
        A1 = np.sum((dOP * dOP.conj()).real + eps, axis=(-2, -1))
        A1 = A1 + 0.5 * np.mean(A1, axis=-3)
        out.update(object_update_precond=precond)
    if recover_probe:
        dPO = m_probe_update[..., 0:1, :, :] * patches
        A4 = np.sum((dPO * dPO.conj()).real + eps, axis=(-2, -1))
        A4 = A4 + 0.5 * np.mean(A4, axis=-3)
    if recover_psi and recover_probe:
        b1 = np.sum((dOP.conj() * chi0).real, axis=(-2, -1))
        b2 = np.sum((dPO.conj() * chi0).real, axis=(-2, -1))
        A2 = np.sum(dOP * dPO.conj(), axis=(-2, -1))
        A3 = A2.conj()
        det = A1 * A4 - A2 * A3
        x1 = -np.conj(A2 * b2 - A4 * b1) / det
        x2 = np.conj(A1 * b2 - A3 * b1) / det
    elif recover_psi:
        b1 = np.sum((dOP.conj() * chi0).real, axis=(-2, -1))
        x1 = b1 / A1
    elif recover_probe:
        b2 = np.sum((dPO.conj() * chi0).real, axis=(-2, -1))
        x2 = b2 / A4
    if recover_psi:
        step = 0.9 * np.maximum(0, x1[..., None, None].real)
        out.update(beta_object=np.mean(step, axis=-5)[..., 0, 0, 0])
    if recover_probe:
        step = 0.9 * np.maximum(0, x2[..., None, None].real)
        out.update(beta_probe=np.mean(step, axis=-5))
    return out


def update_position(scan, pos_num, pos_den, *, alpha=0.05, limit=0.0):
    """lstsq._update_position (lstsq.py:764-806) without Adam."""
    step = pos_num / ((1 - alpha) * pos_den +
                      alpha * max(pos_den.max(), 1e-6))
    if limit > 0:
        step = np.clip(step, -limit, limit)
    step = step - scipy.stats.trim_mean(step, 0.05)
    return (scan - step).astype(f32)


def lstsq_epoch(data, scan, psi, probe, mask, batches, order, *,
                psi_precond=None, recover_psi=True, recover_probe=True,
                recover_positions=False, position_limit=0.0, **kw):
    """One call of solvers.lstsq_grad (lstsq.py:25-294), non-compact batches,
    no adaptive moment, no eigen probes.  Returns psi, probe, scan, cost."""
    if psi_precond is None:
        psi_precond = psi_preconditioner(psi, probe, scan)
    num_batch = len(batches)
    batch_cost = np.zeros(num_batch, dtype=f32)
    pos_num = np.zeros_like(scan)
    pos_den = np.zeros_like(scan)
    probe = probe.copy()
    for n in order:
        b = batches[n]
        lo, hi = b[0], b[-1] + 1
        r = lstsq_batch(data[lo:hi], scan[lo:hi], psi, probe, mask,
                        psi_precond, num_batch, recover_psi=recover_psi,
                        recover_probe=recover_probe,
                        recover_positions=recover_positions, **kw)
        if recover_psi:
            psi = (psi + r['beta_object'] * r['object_update_precond']
                  ).astype(c64)
        if recover_probe:
            probe = (probe + r['beta_probe'] * r['m_probe_update']).astype(c64)
        if recover_positions:
            pos_num[lo:hi] = r['pos_num']
            pos_den[lo:hi] = r['pos_den']
        batch_cost[n] = np.mean(r['costs'])
    if recover_positions:
        scan = update_position(scan, pos_num, pos_den, limit=position_limit)
    return psi, probe, scan, float(batch_cost.mean())


# ----------------------------------------------------------------------------
# Multislice objects, D > 1 (rPIE only in the reference)
# (operators/cupy/multislice.py, fresnelspectprop.py, rpie.py:374, 441-474,
#  _preconditioner.py:48-167)
# ----------------------------------------------------------------------------

def fresnel_propagator(n, probe_fov, distance, wavelength):
    """fresnelspectprop.py:113-135: fftshift(exp(i z sqrt(k^2 - Kx^2 - Ky^2)))
    on the grid (0.5 + linspace(-n/2, n/2 - 1, n)) / n, as complex64."""
    grid = (0.5 + np.linspace(-0.5 * n, 0.5 * n - 1, num=n)) / n
    kx = 2 * np.pi * n * grid / probe_fov[1]
    ky = 2 * np.pi * n * grid / probe_fov[0]
    Kx, Ky = np.meshgrid(kx, ky, indexing='xy')
    h = np.exp(1j * distance * np.sqrt((2 * np.pi / wavelength)**2 - Kx**2 - Ky**2))
    return np.fft.fftshift(h).astype(c64)


def fresnel_fwd(x, h, norm='ortho'):
    """fresnelspectprop.py:52-80."""
    return ifft2(fft2(x, norm) * h, norm)


def fresnel_adj(x, h, norm='ortho'):
    """fresnelspectprop.py:82-111."""
    return ifft2(fft2(x, norm) * np.conj(h), norm)


def multislice_exitwave(psi, scan, probe, h):
    """multislice.py:97-139: exit wave of the last slice (B, M, N, N) and the
    probe incident on every slice (D, B, M, N, N).  probe is (B|1, M, N, N);
    probe width == detector width."""
    D, B = psi.shape[0], len(scan)
    M, N = probe.shape[-3], probe.shape[-1]
    probes = np.zeros((D, B, M, N, N), dtype=c64)
    probes[0] = probe
    for t in range(D):
        ew = exitwave(psi[t], scan, probes[t], N)
        if t == D - 1:
            break
        probes[t + 1] = fresnel_fwd(ew, h)
    return ew, probes


def multislice_farplane(psi, scan, probe, h, norm='ortho'):
    """Ptycho.fwd for D >= 1 slices: (B, 1, M, N, N)."""
    ew, _ = multislice_exitwave(psi, scan, probe[..., 0, :, :, :], h)
    return fft2(ew, norm)[:, None]


def rpie_batch_multislice(data, scan, psi, probe, mask, h, *, eigen_probe=None,
                          eigen_weights=None, noise_model='gaussian',
                          unmeasured_scaling=1.0, norm='ortho', chunk=64,
                          step_length_start=0.5, step_length_weight=0.5,
                          usemodes='all_modes', recover_probe=True):
    """rpie._get_nearplane_gradients (rpie.py:315-567) for a (D, H, W) object.

    Returns costs (B,), psi_numerator (D, H, W), probe_numerator
    (D, 1, 1, M, N, N), eigen_weights.  Reference behaviour kept: the residual
    goes to the previous slice through the adjoint Fresnel step only
    (rpie.py:474)."""
    B, D = len(scan), psi.shape[0]
    M, N = probe.shape[-3], probe.shape[-1]
    psi_numerator = np.zeros_like(psi)
    probe_numerator = np.zeros((D, *probe.shape), dtype=c64)
    costs = np.empty(B, dtype=f32)
    if eigen_weights is not None:
        eigen_weights = eigen_weights.copy()
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        unique = get_varying_probe(
            probe, eigen_probe,
            eigen_weights[lo:hi] if eigen_weights is not None else None)
        ew, probes = multislice_exitwave(psi, scan[lo:hi], unique[:, 0], h)
        far = fft2(ew, norm)[:, None]
        chi_hat, costs[lo:hi] = farplane_gradient(
            far, data[lo:hi], mask, noise_model, unmeasured_scaling,
            step_length_start, step_length_weight, usemodes)
        diff = ifft2(chi_hat, norm)  # (b, 1, M, N, N)
        for t in range(D - 1, -1, -1):
            grad_psi = (np.conj(probes[t][:, None]) * diff / M).reshape(
                (hi - lo) * M, N, N).astype(c64)
            psi_numerator[t] = patch_adj(scan[lo:hi], grad_psi, psi_numerator[t], N,
                                         nrepeat=M)
            patches = patch_fwd(psi[t], scan[lo:hi], N)[:, None, None]
            probe_numerator[t] += np.sum(np.conj(patches) * diff, axis=0, keepdims=True)
            if t == 0:
                break
            diff = fresnel_adj(diff, h, norm)
        if recover_probe and eigen_weights is not None:
            patches = patch_fwd(psi[0], scan[lo:hi], N)[:, None, None]
            OP = patches * probe[..., 0:1, :, :]
            num = np.sum(np.real(np.conj(OP) * diff[..., 0:1, :, :]), axis=(-1, -2))
            den = np.sum(np.abs(OP)**2, axis=(-1, -2))
            eigen_weights[lo:hi, 0:1, 0:1] += f32(0.1) * (num / den)
    return costs, psi_numerator, probe_numerator, eigen_weights


def psi_preconditioner_multislice(psi, probe, scan, h, chunk=64):
    """_preconditioner.py:48-100: slice 0 from the shared probe; slice i from
    the probe propagated through slices 0..i-1 at every position."""
    out = np.zeros(psi.shape, dtype=c64)
    N = probe.shape[-1]
    for lo in range(0, len(scan), chunk):
        sc = scan[lo:lo + chunk]
        amp = np.sum(probe * np.conj(probe), axis=-3)[:, 0]
        out[0] = patch_adj(sc, amp, out[0], N)
        probe1 = probe[:, 0]
        for i in range(1, psi.shape[0]):
            probe1 = fresnel_fwd(exitwave(psi[i - 1], sc, probe1, N), h)
            amp = np.sum(probe1 * np.conj(probe1), axis=-3)
            out[i] = patch_adj(sc, amp.astype(c64), out[i], N)
    return out


def probe_preconditioner_multislice(psi, probe, scan, chunk=64):
    """_preconditioner.py:116-167: one (N, N) plane per slice."""
    N = probe.shape[-1]
    out = np.zeros((psi.shape[0], N, N), dtype=c64)
    for lo in range(0, len(scan), chunk):
        for i in range(psi.shape[0]):
            patches = patch_fwd(psi[i], scan[lo:lo + chunk], N)
            out[i] += np.sum(patches * np.conj(patches), axis=0)
    return out
