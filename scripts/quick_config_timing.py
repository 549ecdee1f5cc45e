"""Epoch timings of the non-headline BASELINE configurations (development aid).
Usage: python scripts/quick_config_timing.py [lstsq256|dm512|lstsq64|rpie_eigen128]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import tike_b200.ptycho as tp  # noqa: E402
from tike_b200 import kernels as K, synthetic  # noqa: E402


def run(name, algo, det, M, P, H, num_batch, epochs=3, eigen=False, position=False,
        noise='gaussian', slices=1, probe_width=None):
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    amp = 0.8 + 0.2 * torch.rand((H, H), device=dev, generator=g)
    psi_true = torch.polar(amp, torch.rand((H, H), device=dev, generator=g) - 0.5).to(torch.complex64)[None].contiguous()
    N = probe_width or det  # probe narrower than the detector: zero-padded exit wave
    probe = synthetic.make_probe(N, M, seed=2, photons=float(det * det) * 50)
    scan = synthetic.make_scan(P, H, H, N, seed=1)
    scan_d = torch.as_tensor(scan, device=dev)
    probe_d = torch.as_tensor(probe[0, 0], device=dev)
    data = torch.empty((P, det, det), dtype=torch.float32, device=dev)
    far = torch.empty((256, M, det, det), dtype=torch.complex64, device=dev) if det > 128 else None
    for lo in range(0, P, 256):
        hi = min(P, lo + 256)
        b = K.make_batch(psi_true[0], scan_d[lo:hi].contiguous(), probe_d, det)
        K.ptycho_fwd(b, far[:hi - lo] if far is not None else None, data[lo:hi])
    del far
    alg = {'rpie': tp.RpieOptions(num_batch=num_batch, alpha=0.2, batch_method='compact'),
           'lstsq_grad': tp.LstsqOptions(num_batch=num_batch, batch_method='compact'),
           'dm': tp.DmOptions(num_batch=num_batch, batch_method='compact')}[algo]
    psi0 = np.full((slices, H, H), 0.5 + 0j, np.complex64)
    if noise == 'poisson':
        # a flat start has exactly-zero far-field pixels, and the reference's
        # Poisson step divides by the intensity without eps (rpie.py:390)
        rng = np.random.default_rng(0)
        psi0 = (psi0 * (1 + 0.2 * rng.standard_normal(psi0.shape)) *
                np.exp(0.3j * rng.standard_normal(psi0.shape))).astype(np.complex64)
    ew = None
    if eigen:
        ew = np.ones((P, 1, M), np.float32)
    params = tp.PtychoParameters(
        probe=probe, scan=scan,
        psi=psi_true.cpu().numpy() if position else psi0,
        eigen_weights=ew, algorithm_options=alg,
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool),
                                            noise_model=noise),
        position_options=tp.PositionOptions(initial_scan=scan.copy(), update_magnitude_limit=0.5) if position else None,
        probe_options=tp.ProbeOptions(probe_wavelength=1.2e-10,
                                      probe_FOV_lengths=(det * 2e-8, det * 2e-8)),
        object_options=tp.ObjectOptions(multislice_propagation_distance=3e-6))
    order = np.arange(P)
    split = ([order], [np.array_split(order, num_batch)], [0])
    with tp.Reconstruction(data, params, split=split, data_is_local=True) as ctx:
        ctx.iterate(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.iterate(epochs)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / epochs
        costs = [c[0] for c in ctx.parameters.algorithm_options.costs]
    print(f'{name}: {algo} det={det} M={M} P={P} batches={num_batch}: {dt*1e3:.1f} ms/epoch '
          f'-> {P/dt:.0f} patterns/s; costs {costs[0]:.3g} -> {costs[-1]:.3g}', flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['lstsq256', 'dm512', 'lstsq64', 'rpie_eigen128', 'lstsq128']
    if 'lstsq256' in which:
        run('config3-like', 'lstsq_grad', 256, 4, 4000, 2048, 2)
    if 'dm512' in which:
        run('config5-like', 'dm', 512, 1, 2000, 4096, 1)
    if 'lstsq64' in which:
        run('config1-like', 'lstsq_grad', 64, 1, 20000, 1024, 2)
    if 'lstsq128' in which:
        run('lstsq 128x8', 'lstsq_grad', 128, 8, 20000, 2048, 2)
    if 'rpie256' in which:
        run('rPIE 256x4', 'rpie', 256, 4, 4000, 2048, 2)
    if 'rpie256poisson' in which:
        run('rPIE 256x4 poisson', 'rpie', 256, 4, 4000, 2048, 2, noise='poisson')
    if 'lstsq128big' in which:
        run('lstsq 128x8 at the bench size', 'lstsq_grad', 128, 8, 100000, 4096, 5, epochs=2)
    if 'rpie128pad' in which:
        run('rPIE 128x8, probe 96', 'rpie', 128, 8, 20000, 2048, 2, probe_width=96)
    if 'rpie128ms2' in which:
        run('rPIE 128x8, 2 slices', 'rpie', 128, 8, 8000, 2048, 2, slices=2)
    if 'lstsq128pos' in which:
        run('lstsq 128x8 + positions', 'lstsq_grad', 128, 8, 20000, 2048, 2, position=True)
    if 'rpie128poisson' in which:
        run('rPIE 128x8 poisson', 'rpie', 128, 8, 20000, 2048, 2, noise='poisson')
    if 'rpie_eigen128' in which:
        run('config4-like', 'rpie', 128, 8, 20000, 2048, 2, eigen=True)
