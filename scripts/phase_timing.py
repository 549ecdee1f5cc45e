"""Per-phase SM cycles of the stage-fused rPIE kernel (development aid).

Needs a library built with  TB_NVCC_EXTRA=-DTB_PHASE_TIMING python -m tike_b200.build --force
(rebuild without the switch afterwards; the timed build is not the product), or a
variant next to the product:  python scripts/build_variant.py phase rpie_p3.cu -DTB_PHASE_TIMING
and TB_LIB_PATH=tike_b200/lib/libtikeb200_phase.so python scripts/phase_timing.py p3
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, '.')
from tike_b200 import kernels as K  # noqa: E402
from tike_b200._lib import lib  # noqa: E402

P3_NAMES = ['patch load', 'pass 1 fwd', 'pass 2 fwd', 'pass 3 fwd+spill', 'cost/modulus',
            'pass 3 inv+reload', 'pass 2 inv', 'pass 1 inv+grad', 'scatter', '-', '-', 'loop head']
NAMES = ['patch load', 'colA fwd', 'rowA fwd', 'rowB fwd', 'colB fwd+spill', 'cost/modulus',
         'colB inv+reload', 'rowB inv', 'rowA inv', 'colA inv+grad', 'scatter', 'loop head']


def main(det=128, M=8, B=148 * 8, H=2048, W=2048, p3=False):
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
    h = lib()
    out = (C.c_ulonglong * 16)()
    dbg = h.tb_debug_phases_p3 if p3 else h.tb_debug_phases
    names = P3_NAMES if p3 else NAMES
    per_mode = (1, 2, 3, 5, 6, 7) if p3 else (1, 2, 3, 4, 6, 7, 8, 9)
    for it in range(2):
        dbg(None, 1)
        K.rpie_batch(b, data, None, det * det, noise_model='gaussian',
                     psi_numerator=psi_num, probe_numerator=probe_num, costs=costs)
        dbg(out, 0)
    tot = sum(out[:12])
    print(f'det={det} M={M} B={B}: cycles per position (mean over CTAs)')
    for i, n in enumerate(names):
        per = out[i] / B
        print(f'  {n:18s} {per:10.0f} cyc  {100.0 * out[i] / tot:5.1f} %'
              + (f'   ({per / M:8.0f} per mode)' if i in per_mode else ''))
    print(f'  total {tot / B:.0f} cycles per position')


if __name__ == '__main__':
    # "p3": the three-pass kernel (rpie_p3.cu); default: rpie_fast_kernel (run with TB_RPIE_P3=0)
    main(p3='p3' in sys.argv[1:])
