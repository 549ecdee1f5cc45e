"""GPU stand-in for the reference's CuPy path (REPORTED BASELINE ONLY).

The reference cannot run on the GPU box (CuPy is not installed and there is no
network), so SURVEY.md §8(d)(i) asks for its op sequence executed 1:1 with
library GPU ops on the same B200.  This module restates, op for op, what
``rpie._get_nearplane_gradients`` launches for one 64-pattern chunk
(src/tike/ptycho/solvers/rpie.py:355-505): bilinear patch gather
(operators/cupy/convolution.cu:79-144 semantics) written M times, ``*= probe``,
cuFFT forward (``torch.fft`` = the same cuFFT as ``cupyx.scipy.fft``),
intensity, Gaussian cost, masked modulus step, cuFFT inverse, ``conj(probe)*chi``
scatter-add with the four bilinear weights, ``conj(patch)*chi`` probe sum.

It is not part of the product (nothing under ``tike_b200/`` imports it) and it
is not the oracle either; ``bench.py`` times it next to the fused kernel and
``tests/test_gpu_kernels.py`` checks that it computes the same numerators.
"""
from __future__ import annotations

import torch

CHUNK = 64  # stream_and_modify2 chunk size (communicators/stream.py:285-404)


def _corners(scan, N):
    iy = torch.floor(scan[:, 0])
    ix = torch.floor(scan[:, 1])
    fy = (scan[:, 0] - iy)[:, None, None]
    fx = (scan[:, 1] - ix)[:, None, None]
    ar = torch.arange(N, device=scan.device)
    yy = iy.long()[:, None, None] + ar[None, :, None]
    xx = ix.long()[:, None, None] + ar[None, None, :]
    w = ((1 - fy) * (1 - fx), (1 - fy) * fx, fy * (1 - fx), fy * fx)
    return yy, xx, w


def rpie_chunk(data, scan, psi, probe, psi_num, probe_num):
    """One chunk; psi (H, W), probe (M, N, N), data (B, N, N).  Accumulates into
    psi_num (H, W) and probe_num (M, N, N); returns the per-pattern costs."""
    M, N = probe.shape[0], probe.shape[-1]
    H, W = psi.shape
    yy, xx, w = _corners(scan, N)
    flat = psi.reshape(-1)
    i00 = yy * W + xx
    patch = (w[0] * flat[i00] + w[1] * flat[i00 + 1] +
             w[2] * flat[i00 + W] + w[3] * flat[i00 + W + 1])
    # Patch.fwd(nrepeat=M) writes the patch once per mode, then *= probe
    near = patch[:, None].repeat(1, M, 1, 1)
    near *= probe[None]
    far = torch.fft.fft2(near, norm='ortho')
    intensity = torch.sum(far.real ** 2 + far.imag ** 2, dim=1)
    sd, sI = torch.sqrt(data), torch.sqrt(intensity)
    costs = torch.mean(torch.square(sI - sd), dim=(-2, -1))
    far = far * (-(1.0 - sd / (sI + 1e-9)))[:, None]
    chi = torch.fft.ifft2(far, norm='ortho')
    grad = torch.sum(torch.conj(probe)[None] * chi, dim=1) / M
    pn = psi_num.reshape(-1)
    pr = torch.view_as_real(pn)
    for wk, off in zip(w, (0, 1, W, W + 1)):
        v = torch.view_as_real((wk * grad).reshape(-1).contiguous())
        pr.index_add_(0, (i00 + off).reshape(-1), v)
    probe_num += torch.sum(torch.conj(patch)[:, None] * chi, dim=0)
    return costs


def rpie_batch(data, scan, psi, probe, chunk=CHUNK):
    """All chunks of one batch; returns (costs, psi numerator, probe numerator)."""
    psi_num = torch.zeros_like(psi)
    probe_num = torch.zeros_like(probe)
    costs = torch.empty(scan.shape[0], dtype=torch.float32, device=psi.device)
    for lo in range(0, scan.shape[0], chunk):
        hi = min(scan.shape[0], lo + chunk)
        costs[lo:hi] = rpie_chunk(data[lo:hi], scan[lo:hi], psi, probe, psi_num, probe_num)
    return costs, psi_num, probe_num
