"""Reconstruct a synthetic ptychography scan with the tike API on B200.

    python examples/reconstruct_synthetic.py                     # one GPU
    torchrun --nproc-per-node 4 examples/reconstruct_synthetic.py  # one process per GPU

The same script runs against the reference by replacing ``tike_b200`` with
``tike`` (and dropping the torch.distributed lines): parameter classes, option
names and result fields are the reference's.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tike_b200.ptycho as tp  # noqa: E402
from tike_b200 import synthetic  # noqa: E402


def main(det=128, modes=4, positions=4000, size=1024, epochs=20):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        torch.distributed.init_process_group('nccl')

    # ground truth, probe, scan (seeded) and the "measured" data
    psi_true, probe, scan = synthetic.make_problem(positions, det, modes, size, size, seed=0)
    data = tp.simulate(detector_shape=det, probe=probe, scan=scan, psi=psi_true)

    parameters = tp.PtychoParameters(
        probe=probe, scan=scan,
        psi=np.full_like(psi_true, 0.5 + 0j),
        algorithm_options=tp.RpieOptions(num_batch=5, num_iter=epochs, alpha=0.2),
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((det, det), bool)),
        probe_options=tp.ProbeOptions(force_orthogonality=True),
        object_options=tp.ObjectOptions(),
    )
    result = tp.reconstruct(data, parameters, num_gpu=world)

    if int(os.environ.get('RANK', '0')) == 0:
        costs = [float(np.mean(c)) for c in result.algorithm_options.costs]
        print('cost per epoch:', ' '.join(f'{c:.4g}' for c in costs))
        print(f'{positions * epochs / sum(result.algorithm_options.times):.0f} patterns/s '
              f'on {world} GPU(s), including the first-epoch set-up')
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
