"""Does the order of the positions inside a batch matter to the fused rPIE
kernel?  Same batch (20k positions spread over a 4096^2 object, like one
wobbly_center batch of the bench), visited in random order and in band-sorted
order (kernels.band_order).  Development aid."""
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
    scan_rand = scan_all[pick].contiguous()
    scan_sort = scan_rand[K.band_order(scan_rand).long()].contiguous()
    B = scan_rand.shape[0]
    data = torch.rand((B, det, det), device=dev, generator=g) * 100
    costs = torch.empty(B, device=dev)
    psi_num = torch.zeros_like(psi)
    probe_num = torch.empty_like(probe)
    for name, scan in (('random', scan_rand), ('band-sorted', scan_sort),
                       ('random', scan_rand), ('band-sorted', scan_sort)):
        b = K.make_batch(psi, scan, probe, det)
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
        best = min(ms[1:])
        print(f'{name:12s} B={B}: {best:.2f} ms -> {B / best * 1e3:.0f} patterns/s  {["%.2f" % m for m in ms]}',
              flush=True)


if __name__ == '__main__':
    main()
