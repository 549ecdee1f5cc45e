"""Quick device-side timing of the fused rPIE batch kernel (development aid)."""
import sys

import torch

sys.path.insert(0, '.')
from tike_b200 import kernels as K  # noqa: E402


def main(det=128, M=8, B=4096, H=2048, W=2048):
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    psi = torch.complex(torch.rand((H, W), device=dev, generator=g) + 0.5,
                        torch.rand((H, W), device=dev, generator=g) - 0.5).contiguous()
    probe = torch.complex(torch.rand((M, det, det), device=dev, generator=g),
                          torch.rand((M, det, det), device=dev, generator=g)).contiguous()
    scan = (torch.rand((B, 2), device=dev, generator=g) * (H - det - 4) + 2).contiguous()
    data = torch.rand((B, det, det), device=dev, generator=g) * 100
    b = K.make_batch(psi, scan, probe, det)
    costs = torch.empty(B, device=dev)
    psi_num = torch.zeros_like(psi)
    probe_num = torch.empty_like(probe)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.rpie_batch(b, data, None, det * det, noise_model='gaussian',
                     psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f'det={det} M={M} B={B}: {ms:.2f} ms  -> {B / ms * 1e3:.0f} patterns/s', flush=True)


if __name__ == '__main__':
    main()
    main(det=64, M=1, B=8192, H=1024, W=1024)
