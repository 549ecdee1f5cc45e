"""Thin Python launchers over the C-ABI (torch CUDA tensors in, tensors out).

Every function enqueues work on torch's current CUDA stream and returns
immediately.  Inputs may be any ``__cuda_array_interface__`` exporter; outputs
are allocated with torch (device memory / streams are torch's job here, the
arithmetic is libtikeb200's).
"""
from __future__ import annotations

import ctypes as C
import threading
import typing

import numpy as np
import torch

from . import _lib
from ._lib import tb_batch, tb_rpie_args, tb_lstsq_args, dev_ptr, stream_ptr, check

# number of libtikeb200 kernel launches issued through this module (bench.py
# reports it as gpu_launches); key = C entry point
LAUNCHES: typing.Dict[str, int] = {}


def _count(name: str, n: int = 1):
    LAUNCHES[name] = LAUNCHES.get(name, 0) + n


def launch_count() -> int:
    return sum(LAUNCHES.values())


NOISE = {'gaussian': 0, 'poisson': 1}
STEP_MODE = {'all_modes': 0, 'dominant_mode': 1}


def as_tensor(x, dtype=None, device=None) -> torch.Tensor:
    """View any device array as a torch tensor (zero copy) or upload a host
    array; ensures C-contiguity."""
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, '__cuda_array_interface__'):
        t = torch.as_tensor(x, device='cuda')
    else:
        t = torch.as_tensor(np.ascontiguousarray(x))
    if device is not None or not t.is_cuda:
        t = t.to(device if device is not None else 'cuda')
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def fft_scales(n: int, norm: str = 'ortho') -> typing.Tuple[float, float]:
    """(forward, inverse) factors of the reference's Propagation operator
    (propagation.py:52-57, exitwave.py:71-80)."""
    if norm == 'ortho':
        return 1.0 / n, 1.0 / n
    if norm == 'forward':
        return 1.0 / (n * n), 1.0
    if norm == 'backward':
        return 1.0, 1.0 / (n * n)
    raise ValueError(f'unknown FFT normalization {norm!r}')


def make_batch(psi2d, scan, probe, detector_width, norm='ortho',
               eigen_probe=None, eigen_weights=None) -> tb_batch:
    """Fill a tb_batch.  psi2d (H, W); scan (B, 2); probe (M, N, N) or
    (B, M, N, N); eigen_probe (E, Me, N, N); eigen_weights (B, E+1, M)."""
    b = tb_batch()
    b.psi = dev_ptr(psi2d, '<c8', 'psi')
    b.height, b.width = int(psi2d.shape[-2]), int(psi2d.shape[-1])
    b.scan = dev_ptr(scan, '<f4', 'scan')
    b.npos = int(scan.shape[0])
    b.probe = dev_ptr(probe, '<c8', 'probe')
    b.nmodes, b.probe_width = int(probe.shape[-3]), int(probe.shape[-1])
    b.probe_per_position = 1 if probe.ndim == 4 and probe.shape[0] == scan.shape[0] and probe.shape[0] != 1 else 0
    if probe.ndim == 4 and not b.probe_per_position and probe.shape[0] != 1:
        raise ValueError('probe leading axis must be 1 or the number of positions')
    if eigen_probe is not None:
        b.eigen_probe = dev_ptr(eigen_probe, '<c8', 'eigen_probe')
        b.neigen, b.eigen_modes = int(eigen_probe.shape[-4]), int(eigen_probe.shape[-3])
    else:
        b.eigen_probe = None
        b.neigen = 0 if eigen_weights is None else int(eigen_weights.shape[-2]) - 1
        b.eigen_modes = 0
    if eigen_weights is not None:
        if eigen_weights.shape[-1] != b.nmodes or eigen_weights.shape[0] != b.npos:
            raise ValueError('eigen_weights must be (positions, eigen + 1, modes)')
        if eigen_probe is None and eigen_weights.shape[-2] != 1:
            # weights of absent eigen probes are ignored by get_varying_probe
            pass
        b.eigen_weights = dev_ptr(eigen_weights, '<f4', 'eigen_weights')
        b.neigen = int(eigen_weights.shape[-2]) - 1
    else:
        b.eigen_weights = None
    b.detector_width = int(detector_width)
    b.fwd_scale, b.inv_scale = fft_scales(int(detector_width), norm)
    # the struct only holds raw pointers: keep the arrays alive with it
    b._refs = (psi2d, scan, probe, eigen_probe, eigen_weights)
    return b


# --------------------------------------------------------------------------
def patch_fwd(images, positions, patches, patch_width, nrepeat=1):
    _count('tb_patch_fwd', 1)
    nimage = int(np.prod(images.shape[:-2])) if images.ndim > 2 else 1
    check(_lib.lib().tb_patch_fwd(
        dev_ptr(images, '<c8', 'images'), dev_ptr(patches, '<c8', 'patches'),
        dev_ptr(positions, '<f4', 'positions'), nimage, int(images.shape[-2]),
        int(images.shape[-1]), int(positions.shape[-2]), int(nrepeat),
        int(patch_width), int(patches.shape[-1]), stream_ptr()), 'Patch.fwd')
    return patches


def patch_adj(images, positions, patches, patch_width, nrepeat=1):
    _count('tb_patch_adj', 1)
    nimage = int(np.prod(images.shape[:-2])) if images.ndim > 2 else 1
    check(_lib.lib().tb_patch_adj(
        dev_ptr(images, '<c8', 'images'), dev_ptr(patches, '<c8', 'patches'),
        dev_ptr(positions, '<f4', 'positions'), nimage, int(images.shape[-2]),
        int(images.shape[-1]), int(positions.shape[-2]), int(nrepeat),
        int(patch_width), int(patches.shape[-1]), int(patches.shape[-3]),
        stream_ptr()), 'Patch.adj')
    return images


def fft2(x, inverse=False, scale=1.0):
    """In-place batched 2-D FFT over the last two axes of a c64 tensor."""
    _count('tb_fft2', 1 if int(x.shape[-1]) <= 128 else 2)
    n = int(x.shape[-1])
    if x.shape[-2] != n:
        raise ValueError(f'waves must be square, not {tuple(x.shape)}')
    batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    check(_lib.lib().tb_fft2(dev_ptr(x, '<c8', 'waves'), batch, n,
                             1 if inverse else 0, float(scale), stream_ptr()),
          'Propagation')
    return x


def ptycho_fwd(batch: tb_batch, farplane=None, intensity=None):
    _count('tb_ptycho_fwd', 1)
    check(_lib.lib().tb_ptycho_fwd(
        C.byref(batch), dev_ptr(farplane, '<c8', 'farplane'),
        dev_ptr(intensity, '<f4', 'intensity'), stream_ptr()), 'Ptycho.fwd')


# Per-thread, per-device cache of scratch tensors so the hot loop does not
# allocate (one Python thread drives one GPU, as in the reference's ThreadPool).
_scratch = threading.local()


def scratch(name: str, nbytes: int, device) -> torch.Tensor:
    key = (name, str(device))
    store = getattr(_scratch, 'store', None)
    if store is None:
        store = _scratch.store = {}
    t = store.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        store[key] = t
    return t


def free_scratch():
    store = getattr(_scratch, 'store', None)
    if store is not None:
        store.clear()


def _mask_args(mask, nd):
    if mask is None:
        return None, nd * nd, None
    m = as_tensor(mask).to(torch.uint8).contiguous()
    return dev_ptr(m, '|u1', 'measured_pixels'), None, m


def data_dtype_code(data) -> int:
    ts = data.__cuda_array_interface__['typestr']
    if ts == '<f4':
        return 0
    if ts == '<u2':
        return 1
    raise ValueError(f'diffraction data must be float32 or uint16 on the device, not {ts}')


def rpie_batch(batch: tb_batch, data, mask_u8, num_measured, *, noise_model,
               step_mode='all_modes', step_length_start=0.5,
               step_length_weight=0.5, unmeasured_scaling=1.0,
               psi_numerator=None, probe_numerator=None, costs=None,
               eigen_weight_step=None, device=None):
    """One fused rPIE batch (rpie._get_nearplane_gradients)."""
    a = tb_rpie_args()
    a.batch = batch
    a.data = dev_ptr(data, ('<f4', '<u2'), 'data')
    a.data_dtype = data_dtype_code(data)
    a.mask = dev_ptr(mask_u8, ('|u1', '|b1'), 'measured_pixels')
    a.num_measured = int(num_measured)
    a.noise_model = NOISE[noise_model]
    a.step_mode = STEP_MODE[step_mode]
    a.step_length_start = float(step_length_start)
    a.step_length_weight = float(step_length_weight)
    a.unmeasured_scaling = float(unmeasured_scaling)
    a.accumulate_object = 1 if psi_numerator is not None else 0
    a.psi_numerator = dev_ptr(psi_numerator, '<c8', 'psi_update_numerator')
    a.probe_numerator = dev_ptr(probe_numerator, '<c8', 'probe_update_numerator')
    a.costs = dev_ptr(costs, '<f4', 'costs')
    a.eigen_weight_step = dev_ptr(eigen_weight_step, '<f4', 'eigen_weight_step')
    need = _lib.lib().tb_rpie_workspace_size(C.byref(a))
    ws = scratch('replicas', need, device if device is not None else data.device) if need else None
    a.workspace = dev_ptr(ws) if ws is not None else None
    a.workspace_bytes = int(need)
    _count('tb_rpie_batch', 2 if a.accumulate_object else 1)
    check(_lib.lib().tb_rpie_batch(C.byref(a), stream_ptr()), 'rpie')


# ----------------------------------------------------------- multislice ----
def fresnel_propagator(n, probe_FOV_lengths, distance, wavelength):
    """(n, n) complex64 Fresnel spectrum kernel, DC at the corner
    (fresnelspectprop.py:113-135): exp(i z sqrt(k^2 - kx^2 - ky^2)) on the
    half-sample-shifted frequency grid k_j = 2 pi (j - n/2 + 1/2) / FOV, then
    fftshift.  Evaluated in float64 on the host like the reference."""
    j = np.arange(n, dtype=np.float64) - 0.5 * n + 0.5
    kx = 2.0 * np.pi * j / float(probe_FOV_lengths[1])
    ky = 2.0 * np.pi * j / float(probe_FOV_lengths[0])
    k2 = (2.0 * np.pi / float(wavelength))**2
    phase = float(distance) * np.sqrt(k2 - kx[None, :]**2 - ky[:, None]**2)
    return np.fft.fftshift(np.exp(1j * phase)).astype(np.complex64)


def multislice_batch(psi, scan, probe, detector_width, norm='ortho',
                     eigen_probe=None, eigen_weights=None) -> tb_batch:
    """tb_batch whose psi pointer addresses all (D, H, W) slices."""
    b = make_batch(psi[0], scan, probe, detector_width, norm, eigen_probe, eigen_weights)
    b.psi = dev_ptr(psi, '<c8', 'psi')
    b._refs = b._refs + (psi,)
    return b


def _ms_workspace(batch, nslices, device):
    need = int(_lib.lib().tb_multislice_workspace_size(C.byref(batch), int(nslices)))
    return scratch('multislice', need, device), need


def multislice_fwd(batch: tb_batch, nslices, propagator, farplane=None, intensity=None,
                   device=None):
    """Forward model through D slices (multislice.py:69-95)."""
    dev = device if device is not None else propagator.device
    ws, need = _ms_workspace(batch, nslices, dev)
    _count('tb_multislice_fwd', 4 * nslices + 2)
    check(_lib.lib().tb_multislice_fwd(
        C.byref(batch), int(nslices), dev_ptr(propagator, '<c8', 'propagator'),
        dev_ptr(farplane, '<c8', 'farplane'), dev_ptr(intensity, '<f4', 'intensity'),
        dev_ptr(ws), need, stream_ptr()), 'multislice')


def multislice_rpie_batch(batch: tb_batch, nslices, propagator, data, mask_u8, num_measured,
                          *, noise_model, step_mode='all_modes', step_length_start=0.5,
                          step_length_weight=0.5, unmeasured_scaling=1.0,
                          psi_numerator=None, probe_numerator=None, costs=None,
                          eigen_weight_step=None, device=None):
    """One rPIE batch through a multislice object (rpie.py:355-505, D > 1).
    psi_numerator (D, H, W) accumulated into; probe_numerator (D, M, N, N)."""
    a = tb_rpie_args()
    a.batch = batch
    a.data = dev_ptr(data, ('<f4', '<u2'), 'data')
    a.data_dtype = data_dtype_code(data)
    a.mask = dev_ptr(mask_u8, ('|u1', '|b1'), 'measured_pixels')
    a.num_measured = int(num_measured)
    a.noise_model = NOISE[noise_model]
    a.step_mode = STEP_MODE[step_mode]
    a.step_length_start = float(step_length_start)
    a.step_length_weight = float(step_length_weight)
    a.unmeasured_scaling = float(unmeasured_scaling)
    a.accumulate_object = 1 if psi_numerator is not None else 0
    a.psi_numerator = dev_ptr(psi_numerator, '<c8', 'psi_update_numerator')
    a.probe_numerator = dev_ptr(probe_numerator, '<c8', 'probe_update_numerator')
    a.costs = dev_ptr(costs, '<f4', 'costs')
    a.eigen_weight_step = dev_ptr(eigen_weight_step, '<f4', 'eigen_weight_step')
    ws, need = _ms_workspace(batch, nslices, device if device is not None else data.device)
    a.workspace = dev_ptr(ws)
    a.workspace_bytes = need
    _count('tb_multislice_rpie_batch', 10 * nslices)
    check(_lib.lib().tb_multislice_rpie_batch(
        C.byref(a), int(nslices), dev_ptr(propagator, '<c8', 'propagator'), stream_ptr()),
        'rpie (multislice)')


def multislice_precond_psi(batch: tb_batch, nslices, propagator, out):
    """out (D, H, W) c64 overwritten (_preconditioner.py:48-100)."""
    ws, need = _ms_workspace(batch, nslices, out.device)
    _count('tb_multislice_precond_psi', 4 * nslices)
    check(_lib.lib().tb_multislice_precond_psi(
        C.byref(batch), int(nslices), dev_ptr(propagator, '<c8', 'propagator'),
        dev_ptr(out, '<c8', 'psi_preconditioner'), dev_ptr(ws), need, stream_ptr()),
        'preconditioner (multislice)')


def lstsq_phase1(batch: tb_batch, data, mask_u8, num_measured, *, noise_model,
                 step_mode='all_modes', step_length_start=0.5,
                 step_length_weight=0.5, unmeasured_scaling=1.0, chi,
                 object_upd_sum=None, probe_upd_sum=None, costs=None,
                 position_num=None, position_den=None, taps=None, device=None,
                 nslices=1, propagator=None):
    """lstsq._get_nearplane_gradients for one piece of a batch.  With
    ``nslices > 1`` (``batch`` from multislice_batch, ``propagator`` the Fresnel
    kernel) the far field comes from the multislice forward model and the
    gradients are those of slice 0, as in the fork (lstsq.py:422-530)."""
    a = tb_lstsq_args()
    a.batch = batch
    a.data = dev_ptr(data, ('<f4', '<u2'), 'data')
    a.data_dtype = data_dtype_code(data)
    a.mask = dev_ptr(mask_u8, ('|u1', '|b1'), 'measured_pixels')
    a.num_measured = int(num_measured)
    a.noise_model = NOISE[noise_model]
    a.step_mode = STEP_MODE[step_mode]
    a.step_length_start = float(step_length_start)
    a.step_length_weight = float(step_length_weight)
    a.unmeasured_scaling = float(unmeasured_scaling)
    a.recover_psi = 1 if object_upd_sum is not None else 0
    a.recover_probe = 1 if probe_upd_sum is not None else 0
    a.recover_positions = 1 if position_num is not None else 0
    a.chi = dev_ptr(chi, '<c8', 'chi')
    a.object_upd_sum = dev_ptr(object_upd_sum, '<c8', 'object_upd_sum')
    a.probe_upd_sum = dev_ptr(probe_upd_sum, '<c8', 'probe_upd_sum')
    a.costs = dev_ptr(costs, '<f4', 'costs')
    a.position_num = dev_ptr(position_num, '<f4', 'position_num')
    a.position_den = dev_ptr(position_den, '<f4', 'position_den')
    if taps is not None:
        for i in range(5):
            a.gradient_taps[i] = float(taps[i])
    if int(nslices) > 1:
        ws, need = _ms_workspace(batch, nslices, device if device is not None else data.device)
        a.workspace = dev_ptr(ws)
        a.workspace_bytes = need
        _count('tb_multislice_lstsq_phase1', 4 * int(nslices) + 5)
        check(_lib.lib().tb_multislice_lstsq_phase1(
            C.byref(a), int(nslices), dev_ptr(propagator, '<c8', 'propagator'), stream_ptr()),
            'lstsq_grad (multislice)')
        return
    need = _lib.lib().tb_lstsq_workspace_size(C.byref(a))
    ws = scratch('replicas', need, device if device is not None else data.device) if need else None
    a.workspace = dev_ptr(ws) if ws is not None else None
    a.workspace_bytes = int(need)
    _count('tb_lstsq_phase1', 2 if a.recover_probe else 1)
    check(_lib.lib().tb_lstsq_phase1(C.byref(a), stream_ptr()), 'lstsq_grad')


def lstsq_phase2(batch: tb_batch, chi, object_update, m_probe_update, mode,
                 eps, out):
    _count('tb_lstsq_phase2', 1)
    check(_lib.lib().tb_lstsq_phase2(
        C.byref(batch), dev_ptr(chi, '<c8', 'chi'),
        dev_ptr(object_update, '<c8', 'object_update'),
        dev_ptr(m_probe_update, '<c8', 'm_probe_update'), int(mode),
        float(eps), dev_ptr(out, '<f4', 'out'), stream_ptr()), 'lstsq_grad')


def lstsq_eigen_pass1(batch: tb_batch, chi, mode, m_probe_update, eigen_probe, c,
                      coefs, weights, weight_index, inv_norm_weights, update,
                      intensity_sums=None):
    """Pass 1 of the fused variable-probe update (csrc/eigen.cu).  eigen_probe
    (E, Me, N, N) or None; weights (P, E+1, M) with ``weight_index`` = flat index
    of weights[first position of the batch, c, mode]."""
    _count('tb_lstsq_eigen_pass1', 1)
    nn = int(batch.probe_width) ** 2
    ep = stride = 0
    if eigen_probe is not None:
        ep = dev_ptr(eigen_probe, '<c8', 'eigen_probe') + int(mode) * nn * 8
        stride = int(eigen_probe.shape[-3]) * nn
    wp = wstride = 0
    if weights is not None:
        wp = dev_ptr(weights, '<f4', 'eigen_weights') + int(weight_index) * 4
        wstride = int(weights.shape[-2]) * int(weights.shape[-1])
    check(_lib.lib().tb_lstsq_eigen_pass1(
        C.byref(batch), dev_ptr(chi, '<c8', 'chi'), int(mode),
        dev_ptr(m_probe_update, '<c8', 'm_probe_update'), ep or None, stride, int(c),
        dev_ptr(coefs, '<c8', 'coefs'), int(coefs.shape[-1]) if coefs is not None else 0,
        wp or None, wstride, dev_ptr(inv_norm_weights, '<f4'), dev_ptr(update, '<c8', 'update'),
        dev_ptr(intensity_sums, '<f4', 'intensity_sums'), stream_ptr()), 'lstsq_grad (eigen)')


def lstsq_eigen_pass2(batch: tb_batch, chi, mode, m_probe_update, eigen_probe, c, coefs,
                      n_out, d_out):
    """Pass 2: per-position numerator / denominator of the eigen-weight step
    with the refreshed eigen probe, and its projection coefficient."""
    _count('tb_lstsq_eigen_pass2', 1)
    nn = int(batch.probe_width) ** 2
    ep = dev_ptr(eigen_probe, '<c8', 'eigen_probe') + int(mode) * nn * 8
    check(_lib.lib().tb_lstsq_eigen_pass2(
        C.byref(batch), dev_ptr(chi, '<c8', 'chi'), int(mode),
        dev_ptr(m_probe_update, '<c8', 'm_probe_update'), ep,
        int(eigen_probe.shape[-3]) * nn, int(c), dev_ptr(coefs, '<c8', 'coefs'),
        int(coefs.shape[-1]) if coefs is not None else 0, dev_ptr(n_out, '<f4'),
        dev_ptr(d_out, '<f4'), stream_ptr()), 'lstsq_grad (eigen)')


def _float_scratch(device, n=4):
    return scratch('floats', 4 * n, device).view(torch.float32)


def max_real(x, out=None):
    """max(Re x) over a c64 tensor as a 1-element device float tensor (values
    are non-negative sums; ``out`` accumulates: out = max(out, ...))."""
    _count('tb_max_real', 1)
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=x.device)
    check(_lib.lib().tb_max_real(dev_ptr(x, '<c8'), int(x.numel()), dev_ptr(out, '<f4'),
                                 stream_ptr()), 'max')
    return out


def rpie_update_psi(psi, numerator, precond, alpha, precond_max=None):
    """psi += numerator / ((1 - alpha) precond + alpha max(precond)) in place;
    ``precond_max`` (device float) replaces the in-kernel maximum."""
    if precond_max is not None:
        _count('tb_rpie_update_psi_given_max', 1)
        check(_lib.lib().tb_rpie_update_psi_given_max(
            dev_ptr(psi, '<c8'), dev_ptr(numerator, '<c8'), dev_ptr(precond, '<c8'),
            int(psi.numel()), float(alpha), dev_ptr(precond_max, '<f4'), stream_ptr()),
            'rpie update')
        return
    _count('tb_rpie_update_psi', 2)
    s = _float_scratch(psi.device)
    check(_lib.lib().tb_rpie_update_psi(
        dev_ptr(psi, '<c8'), dev_ptr(numerator, '<c8'), dev_ptr(precond, '<c8'),
        int(psi.numel()), float(alpha), dev_ptr(s), stream_ptr()), 'rpie update')


def rpie_update_psi_adam(psi, numerator, precond, v, m, alpha, vdecay, mdecay,
                         precond_max=None):
    """rpie._update with adaptive moments and no cost history (rpie.py:233-267)
    on one object slice, in place on psi, v (f32) and m (c64)."""
    if precond_max is None:
        precond_max = max_real(precond)
    _count('tb_rpie_update_psi_adam', 1)
    check(_lib.lib().tb_rpie_update_psi_adam(
        dev_ptr(psi, '<c8'), dev_ptr(numerator, '<c8'), dev_ptr(precond, '<c8'),
        dev_ptr(v, '<f4'), dev_ptr(m, '<c8'), int(psi.numel()), float(alpha),
        float(vdecay), float(mdecay), dev_ptr(precond_max, '<f4'), stream_ptr()),
        'rpie update (adam)')


def momentum_update(psi, direction, m, mdecay, beta_dev):
    """m = mdecay m + (1 - mdecay) beta direction; psi += m (lstsq.py:176-193)."""
    _count('tb_momentum_update', 1)
    check(_lib.lib().tb_momentum_update(
        dev_ptr(psi, '<c8'), dev_ptr(direction, '<c8'), dev_ptr(m, '<c8'),
        int(psi.numel()), float(mdecay), dev_ptr(beta_dev, '<f4'), stream_ptr()),
        'momentum update')


def add_quotient(y, numerator, precond, eps, period=0):
    """y += numerator / (Re precond + eps); precond repeats with ``period``."""
    _count('tb_add_quotient', 1)
    check(_lib.lib().tb_add_quotient(
        dev_ptr(y, '<c8'), dev_ptr(numerator, '<c8'), dev_ptr(precond, '<c8'),
        int(y.numel()), int(period), float(eps), stream_ptr()), 'add quotient')


def object_pointwise_constraints(psi, positivity=0.0, clip=False, a_max=1.0):
    _count('tb_object_pointwise_constraints', 1)
    check(_lib.lib().tb_object_pointwise_constraints(
        dev_ptr(psi, '<c8'), int(psi.numel()), float(positivity), 1 if clip else 0,
        float(a_max), stream_ptr()), 'object constraints')


def object_smoothness(psi, a):
    """3x3 smoothing of every (H, W) slice (object.py:227-253); returns a new tensor."""
    _count('tb_object_smoothness', 1)
    out = torch.empty_like(psi)
    lead = int(np.prod(psi.shape[:-2])) if psi.ndim > 2 else 1
    check(_lib.lib().tb_object_smoothness(
        dev_ptr(out, '<c8'), dev_ptr(psi, '<c8'), lead, int(psi.shape[-2]),
        int(psi.shape[-1]), float(a), stream_ptr()), 'object smoothness')
    return out


def weighted_norm_sums(psi, weight):
    """(sum |psi|^2 Re w, sum (Re w)^2) as a float64 device tensor."""
    _count('tb_weighted_norm_sums', 1)
    out = torch.empty(2, dtype=torch.float64, device=psi.device)
    check(_lib.lib().tb_weighted_norm_sums(
        dev_ptr(psi, '<c8'), dev_ptr(weight, '<c8'), int(psi.numel()),
        dev_ptr(out, '<f8'), stream_ptr()), 'weighted norm')
    return out


def scale_by_device_scalar(y, s, divide=False):
    _count('tb_scale_by_device_scalar', 1)
    check(_lib.lib().tb_scale_by_device_scalar(
        dev_ptr(y, '<c8'), int(y.numel()), dev_ptr(s, '<f4'), 1 if divide else 0,
        stream_ptr()), 'scale')


def rpie_update_probe(probe, numerator, probe_precond, alpha):
    _count('tb_rpie_update_probe', 2)
    s = _float_scratch(probe.device)
    n2 = int(probe.shape[-1] * probe.shape[-2])
    check(_lib.lib().tb_rpie_update_probe(
        dev_ptr(probe, '<c8'), dev_ptr(numerator, '<c8'),
        dev_ptr(probe_precond, '<c8'), int(probe.numel() // n2), n2,
        float(alpha), dev_ptr(s), stream_ptr()), 'rpie update')


PRECOND_BAND = 16  # rows per band of band_order(); kBand in csrc/precond.cu


def band_order(scan):
    """Visiting order for the preconditioner kernels: positions sorted by
    (floor(row) // PRECOND_BAND, floor(column)), int32 on the device of
    ``scan``.  The sums do not depend on it; consecutive footprints overlap,
    which is what the window kernels of csrc/precond.cu exploit."""
    import torch
    if not isinstance(scan, torch.Tensor):
        # CuPy / NumPy callers: zero-copy view through the array interface
        scan = torch.as_tensor(scan)
    if scan.shape[0] == 0:
        return torch.empty(0, dtype=torch.int32, device=scan.device)
    corner = torch.floor(scan).to(torch.int64)
    band = torch.div(corner[:, 0], PRECOND_BAND, rounding_mode='floor')
    # no host synchronisation: columns are offset / clamped into 21 bits
    col = (corner[:, 1] + (1 << 20)).clamp_(0, (1 << 21) - 1)
    key = band * (1 << 21) + col
    return torch.argsort(key).to(torch.int32)


def precond_psi(probe, scan, out, order=None):
    """probe (M, N, N), scan (P, 2), out (H, W) c64 overwritten; ``order`` is
    an optional visiting order (see band_order)."""
    _count('tb_precond_psi', 2)
    n = int(probe.shape[-1])
    s = scratch('probe_amp', 4 * n * n, out.device)
    check(_lib.lib().tb_precond_psi(
        dev_ptr(probe, '<c8'), int(probe.shape[-3]), n, dev_ptr(scan, '<f4'),
        _order_ptr(order, scan), int(scan.shape[0]), dev_ptr(out, '<c8'),
        int(out.shape[-2]), int(out.shape[-1]), dev_ptr(s), stream_ptr()),
        'preconditioner')


def precond_probe(psi2d, scan, out, order=None):
    """psi2d (H, W), scan (P, 2), out (N, N) c64 overwritten; ``order`` as in
    precond_psi."""
    _count('tb_precond_probe', 1)
    check(_lib.lib().tb_precond_probe(
        dev_ptr(psi2d, '<c8'), int(psi2d.shape[-2]), int(psi2d.shape[-1]),
        dev_ptr(scan, '<f4'), _order_ptr(order, scan), int(scan.shape[0]),
        int(out.shape[-1]), dev_ptr(out, '<c8'), stream_ptr()), 'preconditioner')


def _order_ptr(order, scan):
    if order is None:
        return None
    if int(order.shape[0]) != int(scan.shape[0]):
        raise ValueError('order must name every position once: '
                         f'{tuple(order.shape)} for {int(scan.shape[0])} positions')
    return dev_ptr(order, '<i4') if order.shape[0] else None


def lstsq_precondition_object(out, upd, precond, alpha=0.05, precond_max=None):
    if precond_max is not None:
        _count('tb_lstsq_precondition_object_given_max', 1)
        check(_lib.lib().tb_lstsq_precondition_object_given_max(
            dev_ptr(out, '<c8'), dev_ptr(upd, '<c8'), dev_ptr(precond, '<c8'),
            int(out.numel()), float(alpha), dev_ptr(precond_max, '<f4'), stream_ptr()),
            'lstsq precondition')
        return
    _count('tb_lstsq_precondition_object', 2)
    s = _float_scratch(out.device)
    check(_lib.lib().tb_lstsq_precondition_object(
        dev_ptr(out, '<c8'), dev_ptr(upd, '<c8'), dev_ptr(precond, '<c8'),
        int(out.numel()), float(alpha), dev_ptr(s), stream_ptr()),
        'lstsq precondition')


def caxpy(y, x, a=1.0, a_dev=None):
    _count('tb_caxpy', 1)
    check(_lib.lib().tb_caxpy(dev_ptr(y, '<c8'), dev_ptr(x, '<c8'),
                              int(y.numel()), float(a),
                              dev_ptr(a_dev, '<f4') if a_dev is not None else None,
                              stream_ptr()), 'caxpy')
