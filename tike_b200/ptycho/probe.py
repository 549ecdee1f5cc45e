"""Probe options, the varying-probe model and per-epoch probe constraints
(reference: src/tike/ptycho/probe.py).

Functions that run every epoch inside a reconstruction take and return torch
CUDA tensors; initialisation helpers (``add_modes_*``, ``init_varying_probe``,
``gaussian``) are host-side NumPy like in the reference.
"""
from __future__ import annotations

import dataclasses
import logging
import typing

import numpy as np
import torch

from .. import linalg, precision
from .. import random as tb_random
from .._array import to_device, to_host

logger = logging.getLogger(__name__)


@dataclasses.dataclass
class ProbeOptions:
    """Settings and state of the probe update (probe.py:55-165)."""

    update_start: int = 0
    update_period: int = 1
    init_rescale_from_measurements: bool = True
    probe_photons: float = np.nan
    probe_wavelength: float = np.nan
    probe_FOV_lengths: typing.Tuple[float, float] = (np.nan, np.nan)
    force_orthogonality: bool = False
    force_centered_intensity: bool = False
    force_sparsity: float = 0.0
    use_adaptive_moment: bool = False
    vdecay: float = 0.999
    mdecay: float = 0.9
    v: typing.Any = dataclasses.field(init=False, default=None)
    m: typing.Any = dataclasses.field(init=False, default=None)
    probe_support: float = 0.0
    probe_support_radius: float = 0.5 * 0.7
    probe_support_degree: float = 2.5
    additional_probe_penalty: float = 0.0
    median_filter_abs_probe: bool = False
    median_filter_abs_probe_px: typing.Tuple[float, float] = (1.0, 1.0)
    preconditioner: typing.Any = dataclasses.field(init=False, default=None)
    power: typing.List[typing.List[float]] = dataclasses.field(
        init=False, default_factory=list)

    def recover_probe(self, epoch: int) -> bool:
        """Whether the probe is updated at this epoch (probe.py:167-169)."""
        return (epoch >= self.update_start) and (epoch % self.update_period == 0)

    def _clone(self) -> "ProbeOptions":
        init_fields = {
            f.name: getattr(self, f.name)
            for f in dataclasses.fields(self) if f.init
        }
        o = ProbeOptions(**init_fields)
        o.power = self.power
        return o

    def copy_to_device(self) -> "ProbeOptions":
        o = self._clone()
        o.v, o.m = to_device(self.v), to_device(self.m)
        o.preconditioner = to_device(self.preconditioner, dtype='c64')
        return o

    def copy_to_host(self) -> "ProbeOptions":
        o = self._clone()
        o.v, o.m = to_host(self.v), to_host(self.m)
        o.preconditioner = to_host(self.preconditioner)
        # entries appended during a reconstruction stay on the device until
        # somebody looks at them (no host synchronisation per epoch)
        o.power = [to_host(x) for x in self.power]
        self.power[:] = o.power
        return o

    def resample(self, factor: float, interp) -> "ProbeOptions":
        return self._clone()  # momentum restarts when the grid changes


# ---------------------------------------------------------------- device ----
def get_varying_probe(shared_probe, eigen_probe=None, weights=None):
    """w0 * probe + sum_c w_c * eigen_c per position (probe.py:272-303).

    shared_probe (1, 1, M, N, N); eigen_probe (1, E, Me, N, N);
    weights (B, E+1, M) -> (B, 1, M, N, N).  The fused kernels evaluate this
    on the fly; this materialising version serves the operator seam."""
    if weights is None:
        return shared_probe.clone()
    unique = weights[..., [0], :, None, None] * shared_probe
    if eigen_probe is not None:
        m = eigen_probe.shape[-3]
        for c in range(eigen_probe.shape[-4]):
            unique[..., :m, :, :] += (weights[..., [c + 1], :m, None, None] *
                                      eigen_probe[..., [c], :m, :, :])
    return unique


def power(probe):
    """Power of each probe mode (probe.py:773-781)."""
    return torch.square(linalg.norm(probe, axis=(-2, -1))).flatten()


def orthogonalize_eig(x):
    """Orthogonalise modes with the eigenvectors of the mode Gram matrix,
    sorted by decreasing power (probe.py:726-770).  NumPy in, NumPy out (the
    reference's set-up code calls it on host arrays); tensors stay tensors."""
    if isinstance(x, np.ndarray):
        modes, power = orthogonalize_eig(torch.from_numpy(np.ascontiguousarray(x)))
        return modes.numpy(), power.numpy()
    flat = x.reshape(*x.shape[:-2], -1)
    # upper triangle of x^H x, like the reference (UPLO='U')
    A = torch.einsum('...ip,...jp->...ij', flat.conj(), flat)
    _, vectors = torch.linalg.eigh(A, UPLO='U')
    result = (vectors.transpose(-1, -2) @ flat).reshape(x.shape)
    pw = torch.square(linalg.norm(result, axis=(-2, -1))).flatten()
    order = torch.flip(torch.argsort(pw, stable=True), dims=(0,))
    return result[..., order, :, :], pw[order]


def finite_probe_support(probe, *, radius=0.5, degree=5.0, p=1.0):
    """Super-Gaussian penalty mask (probe.py:937-981)."""
    if p <= 0:
        return 0.0
    N = probe.shape[-1]
    centers = torch.linspace(-0.5, 0.5 - 1.0 / N, N, dtype=torch.float64,
                             device=probe.device) + 0.5 / N
    i, j = torch.meshgrid(centers, centers, indexing='xy')
    mask = 1 - torch.exp(-(torch.square(i / radius) +
                           torch.square(j / radius))**degree)
    return p * mask.to(torch.float32)


def rescale_probe_using_fixed_intensity_photons(probe, Nphotons,
                                                probe_power_fraction=None):
    """Scale shared modes so their intensities add up to Nphotons
    (probe.py:984-1013)."""
    photons = torch.sum(probe.abs()**2, dim=(-1, -2))
    if probe_power_fraction is None:
        probe_power_fraction = photons / torch.sum(photons)
    return probe * torch.sqrt(probe_power_fraction * Nphotons /
                              photons)[..., None, None]


def constrain_variable_probe(variable_probe, weights):
    """Normalise, orthogonalise and sort eigen probes; clip outlier weights
    (probe.py:306-359)."""
    vnorm = linalg.mnorm(variable_probe, axis=(-2, -1), keepdims=True)
    variable_probe = variable_probe / vnorm
    pwm = variable_probe.shape[-3]
    weights = weights.clone()
    weights[..., 1:, :pwm] *= vnorm[..., 0, 0]
    variable_probe = linalg.orthogonalize_gs(variable_probe, axis=(-2, -1), N=-4)
    pw = linalg.norm(weights[..., 1:, :pwm], keepdims=True, axis=-3)**2
    for i in range(pwm):
        order = torch.argsort(-pw[..., i].flatten())
        weights[..., 1:, i] = weights[..., 1 + order, i]
        variable_probe[..., :, i, :, :] = variable_probe[..., order, i, :, :]
    aevol = weights.abs()
    limit = 1.5 * torch.quantile(aevol, 0.95, dim=-3, keepdim=True)
    weights = torch.minimum(aevol, limit.to(weights.dtype)) * torch.sign(weights)
    return variable_probe, weights


def update_eigen_probe(R, eigen_probe, weights, patches, diff, lo, hi, *,
                       beta=0.1, c=1, m=0, comm=None):
    """Eigen-probe power-iteration-like update (probe.py:362-476).

    R, patches (B,1,1,N,N); diff (B,1,M,N,N); eigen_probe (1,E,Me,N,N);
    weights (P,E+1,M); [lo, hi) is the batch range inside weights.

    With a multi-rank ``comm`` every batch-wide mean runs over the union
    batch of all ranks (sums and counts are all-reduced), so the replicated
    ``eigen_probe`` stays identical on every rank."""

    def union_sum(x):
        """(sum over this rank's positions of x, count) reduced over ranks"""
        total = torch.sum(x, dim=0, keepdim=True)
        count = torch.tensor(float(x.shape[0]), device=x.device)
        if comm is not None and comm.size > 1:
            comm.allreduce_sum_(total)
            comm.allreduce_sum_(count)
        return total, count

    w = weights[lo:hi, c:c + 1, m:m + 1, None, None]
    norm_weights, _ = union_sum(torch.square(w))
    if bool(torch.all(norm_weights == 0)):
        raise ValueError("eigen_probe weights cannot all be zero?")
    ep = eigen_probe[:, c - 1:c, m:m + 1, :, :]
    proj = ((R.conj() * ep).real + w) / norm_weights
    total, count = union_sum(R * torch.mean(proj, dim=(-2, -1), keepdim=True))
    update = (total / count)[0]
    ep = ep + beta * update / linalg.mnorm(update, axis=(-2, -1), keepdims=True)
    ep = ep / linalg.mnorm(ep, axis=(-2, -1), keepdims=True)
    eigen_probe[:, c - 1:c, m:m + 1, :, :] = ep
    phi = patches * ep
    n = torch.mean((diff[:, :, m:m + 1, :, :] * phi.conj()).real, dim=(-1, -2))
    d = torch.mean(torch.square(phi.abs()), dim=(-1, -2))
    total, count = union_sum(d)
    d_mean = (total / count)[0]
    weight_update = (n / (d + 0.1 * d_mean)).reshape(
        weights[lo:hi, c:c + 1, m:m + 1].shape)
    weights[lo:hi, c:c + 1, m:m + 1] += weight_update
    return eigen_probe, weights


def constrain_center_peak(probe):
    """Shift the probes by at most one pixel so the smoothed intensity peak
    moves toward the centre (probe.py:817-861)."""
    import scipy.ndimage
    host = to_host(probe)
    half = host.shape[-2] // 2, host.shape[-1] // 2
    stack = host.reshape((-1, *host.shape[-2:]))
    intensity = scipy.ndimage.gaussian_filter(
        input=np.sum(np.square(np.abs(stack)), axis=0),
        sigma=(half[0] / 3, half[1] / 3), mode="constant", cval=0.0,
        truncate=6.0)
    coords = np.round(scipy.ndimage.center_of_mass(intensity))
    shift = (0, min(1, max(-1, half[0] - coords[0])),
             min(1, max(-1, half[1] - coords[1])))
    shifted = (scipy.ndimage.shift(stack.real, shift, mode="constant", cval=0.0, order=0) +
               1j * scipy.ndimage.shift(stack.imag, shift, mode="constant", cval=0.0, order=0))
    return to_device(shifted.reshape(host.shape).astype(np.complex64),
                     device=probe.device)


def constrain_probe_sparsity(probe, f):
    """Zero the fraction f of pixels with the smallest smoothed intensity
    (probe.py:898-920)."""
    if f == 0:
        return probe
    import scipy.ndimage
    host = to_host(probe).copy()
    stack = host.reshape((-1, *host.shape[-2:]))
    intensity = np.sum(np.square(np.abs(stack)), axis=0)
    sigma = host.shape[-2] / 8, host.shape[-1] / 8
    intensity = scipy.ndimage.gaussian_filter(intensity, sigma=sigma, mode='wrap')
    k = int(f * host.shape[-1] * host.shape[-2])
    smallest = np.argpartition(intensity, k, axis=None)[:k]
    coords = np.unravel_index(smallest, host.shape[-2:])
    host[..., coords[0], coords[1]] = 0
    return to_device(host, device=probe.device)


def apply_median_filter_abs_probe(probe, med_filt_px):
    """Median-filter |probe| of every shared mode (probe.py:864-895)."""
    import scipy.ndimage
    host = to_host(probe).copy()
    mag = scipy.ndimage.median_filter(np.abs(host[0, 0]),
                                      size=(1.0, *med_filt_px), mode="constant")
    host[0, 0] = mag * np.exp(1j * np.angle(host[0, 0]))
    return to_device(host.astype(np.complex64), device=probe.device)


# ------------------------------------------------------------------ host ----
def adjust_probe_power(probe, power=None):
    """Rescale mode powers, default 1/m (probe.py:479-497)."""
    if power is None:
        power = 1.0 / np.arange(1, probe.shape[-3] + 1)
    power = power[..., None, None]
    nrm = linalg.norm(probe, axis=(-2, -1), keepdims=True)
    probe *= power * nrm[..., 0:1, :, :] / nrm
    return probe


def add_modes_random_phase(probe, nmodes):
    """New modes = first mode with random linear phase ramps
    (probe.py:500-531); uses NumPy's legacy global generator."""
    all_modes = np.empty_like(probe, shape=(*probe.shape[:-3], nmodes,
                                            *probe.shape[-2:]))
    pw = probe.shape[-1]
    for m in range(nmodes):
        if m < probe.shape[-3]:
            all_modes[..., m, :, :] = probe[..., m, :, :]
        else:
            shift = np.exp(-2j * np.pi * (np.random.rand(2, 1) - 0.5) *
                           ((np.arange(0, pw) + 0.5) / pw - 0.5))
            all_modes[..., m, :, :] = (probe[..., 0, :, :] * shift[0][None] *
                                       shift[1][:, None])
    return all_modes


def add_modes_cartesian_hermite(probe, nmodes: int):
    """Orthonormal higher modes from Cartesian Hermite-like polynomials times
    the first mode (probe.py:534-644)."""
    if nmodes < 1:
        raise ValueError(f"nmodes cannot be less than 1. It was {nmodes}.")
    if probe.ndim < 3:
        raise ValueError("probe should have shape (..., 1, W, H) "
                         f"not {probe.shape}.")
    ncol = int(np.ceil(np.sqrt(nmodes)))
    nrow = int(np.ceil(nmodes / ncol))
    off = probe.shape[-2] // 2 - 1
    X, Y = np.meshgrid(np.arange(probe.shape[-2]) - off,
                       np.arange(probe.shape[-1]) - off, indexing='xy')
    weight = np.abs(probe)**2
    total = np.sum(weight, axis=(-2, -1), keepdims=True)

    def moment(f):
        return np.sum(f * weight, axis=(-2, -1), keepdims=True) / total

    cenx, ceny = moment(X), moment(Y)
    varx, vary = moment((X - cenx)**2), moment((Y - ceny)**2)
    envelope = np.exp(-((X - cenx)**2 / (2 * varx)) - ((Y - ceny)**2 / (2 * vary)))
    found = []
    for nii in range(nrow):
        for mii in range(ncol):
            basis = ((X - cenx)**mii) * ((Y - ceny)**nii) * probe
            if mii or nii:
                basis = basis * envelope
            basis = basis / linalg.norm(basis, axis=(-2, -1), keepdims=True)
            for H in found:
                basis = basis - H * linalg.inner(H, basis, axis=(-2, -1),
                                                 keepdims=True)
            basis = basis / linalg.norm(basis, axis=(-2, -1), keepdims=True)
            found.append(basis)
            if len(found) == nmodes:
                return np.concatenate(found, axis=-3)[..., :nmodes, :, :]
    raise RuntimeError("add_modes_cartesian_hermite produced too few modes")


def init_varying_probe(scan, shared_probe, num_eigen_probes,
                       probes_with_modes=1):
    """Initial eigen probes and weights (probe.py:666-723)."""
    probes_with_modes = max(probes_with_modes, 0)
    if probes_with_modes > shared_probe.shape[-3]:
        raise ValueError(
            f"probes_with_modes ({probes_with_modes}) cannot be more than "
            f"the number of probes ({shared_probe.shape[-3]})!")
    if num_eigen_probes < 1:
        return None, None
    weights = 1e-6 * np.random.rand(*scan.shape[:-1], num_eigen_probes,
                                    shared_probe.shape[-3]).astype(precision.floating)
    weights -= np.mean(weights, axis=-3, keepdims=True)
    weights[..., 0, :] = 1.0
    weights[..., 1:, probes_with_modes:] = 0
    if num_eigen_probes == 1:
        return None, weights
    eigen_probe = tb_random.numpy_complex(*shared_probe.shape[:-4],
                                          num_eigen_probes - 1,
                                          probes_with_modes,
                                          *shared_probe.shape[-2:])
    eigen_probe /= linalg.mnorm(eigen_probe, axis=(-2, -1), keepdims=True)
    return eigen_probe, weights


def gaussian(size, rin=0.8, rout=1.0):
    """Flat-top probe amplitude with a linear taper (probe.py:784-814)."""
    r, c = np.mgrid[:size, :size] + 0.5
    rs = np.sqrt((r - size / 2)**2 + (c - size / 2)**2)
    rmax = np.sqrt(2) * 0.5 * rout * rs.max() + 1.0
    rmin = np.sqrt(2) * 0.5 * rin * rs.max()
    img = np.zeros((size, size), dtype=precision.floating)
    img[rs < rmin] = 1.0
    zone = np.logical_and(rs > rmin, rs < rmax)
    img[zone] = np.divide(rmax - rs[zone], rmax - rmin)
    return img


def simulate_varying_weights(scan, eigen_probe):
    """Random sinusoidal weights for simulating a varying probe
    (probe.py:647-657): unit amplitude, one sinusoid per eigen probe and mode
    with a period of at most one scan and a random phase; drawn from NumPy's
    legacy global generator like the reference."""
    count = scan.shape[1]
    lead = eigen_probe.shape[:-2]
    step = np.arange(count)[..., :, None, None]
    period = count * np.random.rand(*lead)
    phase = 2 * np.pi * np.random.rand(*lead)
    return np.sin(2 * np.pi / period * step - phase)
