"""Fused rPIE kernel at the bench batch size with float32 and with uint16
patterns, plus the pinned host -> device copy rate of both (development aid:
why is the uint16 end-to-end run not faster than the float32 one?)."""
import sys

import torch

sys.path.insert(0, '.')
from tike_b200 import kernels as K, synthetic  # noqa: E402


def main(det=128, M=8, P=100000, nbatch=5, H=4096):
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    psi = torch.polar(0.8 + 0.2 * torch.rand((H, H), device=dev, generator=g),
                      torch.rand((H, H), device=dev, generator=g) - 0.5).to(torch.complex64).contiguous()
    probe = torch.as_tensor(synthetic.make_probe(det, M, seed=2)[0, 0], device=dev)
    scan_all = torch.as_tensor(synthetic.make_scan(P, H, H, det, seed=1), device=dev)
    pick = torch.randperm(P, device=dev, generator=g)[:P // nbatch]
    scan = scan_all[pick].contiguous()
    scan = scan[K.band_order(scan).long()].contiguous()
    B = scan.shape[0]
    data32 = torch.rand((B, det, det), device=dev, generator=g) * 100
    data16 = torch.round(data32).to(torch.uint16)
    costs = torch.empty(B, device=dev)
    psi_num = torch.zeros_like(psi)
    probe_num = torch.empty_like(probe)
    b = K.make_batch(psi, scan, probe, det)
    for name, data in (('float32', data32), ('uint16', data16), ('float32', data32),
                       ('uint16', data16)):
        ms = []
        for it in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.rpie_batch(b, data, None, det * det, noise_model='gaussian',
                         psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        print(f'kernel {name:8s} B={B}: {min(ms[1:]):.2f} ms', flush=True)
    for name, data in (('float32', data32), ('uint16', data16)):
        host = torch.empty(data.shape, dtype=data.dtype, pin_memory=True)
        host.copy_(data)
        dst = torch.empty_like(data)
        for chunk in (2048, B):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for lo in range(0, B, chunk):
                dst[lo:lo + chunk].copy_(host[lo:lo + chunk], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            print(f'h2d {name:8s} chunks of {chunk:5d}: {t:.2f} ms, '
                  f'{host.numel() * host.element_size() / t / 1e6:.1f} GB/s', flush=True)


if __name__ == '__main__':
    main()
