"""One launch of the fused rPIE kernel exactly as bench.py issues it (20 000
band-sorted positions of a 100k-position scan over a 4096^2 object, 128x128
detector, 8 modes), for `ncu --set full` at the benchmarked size:

    ncu --set full --clock-control none --import-source on -k regex:rpie_fast \\
        -s 2 -c 1 -o gpurun_out/rpie_fast_bench python scripts/profile_rpie_bench.py
"""
import sys

import torch

sys.path.insert(0, '.')
from tike_b200 import kernels as K, synthetic  # noqa: E402


def main(det=128, M=8, P=100000, nbatch=5, H=4096, launches=3):
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
    data = torch.rand((B, det, det), device=dev, generator=g) * 100
    costs = torch.empty(B, device=dev)
    psi_num = torch.zeros_like(psi)
    probe_num = torch.empty_like(probe)
    b = K.make_batch(psi, scan, probe, det)
    for _ in range(launches):
        K.rpie_batch(b, data, None, det * det, noise_model='gaussian',
                     psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
    torch.cuda.synchronize()
    print('launched', launches, 'x', B, 'positions')


if __name__ == '__main__':
    main()
