"""Single launch of the fused rPIE kernel at config-2 tile size, for ncu."""
import sys
import torch
sys.path.insert(0, '.')
from tike_b200 import kernels as K  # noqa: E402

det, M, B, H, W = 128, 8, 148 * 4, 2048, 2048
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
for _ in range(2):
    K.rpie_batch(b, data, None, det * det, noise_model='gaussian',
                 psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
torch.cuda.synchronize()
