"""Where one rPIE epoch of the bench workload goes (development aid)."""
import sys

import torch

sys.path.insert(0, '.')
from tike_b200 import kernels as K, synthetic  # noqa: E402


def timed(f, n=3):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main(P=100000, N=128, M=8, H=4096):
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    psi = torch.polar(0.8 + 0.2 * torch.rand((H, H), device=dev, generator=g),
                      torch.rand((H, H), device=dev, generator=g) - 0.5).to(torch.complex64).contiguous()
    probe = torch.as_tensor(synthetic.make_probe(N, M, seed=2)[0, 0], device=dev)
    scan = torch.as_tensor(synthetic.make_scan(P, H, H, N, seed=1), device=dev)
    pre = torch.empty_like(psi)
    qre = torch.empty((N, N), dtype=torch.complex64, device=dev)
    num = torch.zeros_like(psi)
    print(f'band_order    {timed(lambda: K.band_order(scan)):.3f} ms')
    order = K.band_order(scan)
    shuffled = torch.randperm(P, device=dev, generator=g).to(torch.int32)
    print(f'precond_psi   window {timed(lambda: K.precond_psi(probe, scan, pre, order=order)):.3f} ms'
          f' | direct, sorted {timed(lambda: K.precond_psi(probe, scan, pre)):.3f} ms'
          f' | window, shuffled {timed(lambda: K.precond_psi(probe, scan, pre, order=shuffled)):.3f} ms')
    scan_shuffled = scan[shuffled.long()].contiguous()
    print(f'precond_psi   direct, shuffled scan {timed(lambda: K.precond_psi(probe, scan_shuffled, pre)):.3f} ms')
    print(f'precond_probe window {timed(lambda: K.precond_probe(psi, scan, qre, order=order)):.3f} ms'
          f' | direct, sorted {timed(lambda: K.precond_probe(psi, scan, qre)):.3f} ms'
          f' | direct, shuffled scan {timed(lambda: K.precond_probe(psi, scan_shuffled, qre)):.3f} ms')
    print(f'update_psi    {timed(lambda: K.rpie_update_psi(psi, num, pre, 0.2)):.3f} ms')
    pn = torch.zeros_like(probe)
    print(f'update_probe  {timed(lambda: K.rpie_update_probe(probe, pn, qre, 0.2)):.3f} ms')
    print(f'zeros_like    {timed(lambda: torch.zeros_like(psi)):.3f} ms')


if __name__ == '__main__':
    main()
