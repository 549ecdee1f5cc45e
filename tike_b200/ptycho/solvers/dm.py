"""Difference-map-style solver.

The mounted reference snapshot has NO dm.py and no DmOptions (SURVEY.md §0
F1: only a string compare survives at ptycho/ptycho.py:833), so parity for
this solver is UNPINNED.  Following the survey's design note, ``dm`` is the
rPIE batch pipeline with the object AND probe numerators accumulated over
*all* batches and one update per epoch:

    psi   += sum_batches G_O / (L_O + eps)
    probe += sum_batches G_P / (L_P + eps)

It reuses the fused kernel, is exempt from ``remove_object_ambiguity`` (the
surviving check in the reference) and keeps no per-position exit waves (the
textbook difference map would need P x M x N^2 x 8 bytes resident).
Multislice objects (D > 1) take the per-slice gradients of the rPIE slice
loop (rpie.py:441-474); like there only slice 0 of the probe numerator and
preconditioner drives the probe.
"""
from __future__ import annotations

import torch

from ... import kernels
from ._common import BatchStager, MaskInfo, ObjectReducer, allreduce_
from .rpie import _get_nearplane_gradients


def dm(parameters, data, batches, streams=None, worker_index=0, *, op, epoch,
       comm=None):
    scan, psi, probe = parameters.scan, parameters.psi, parameters.probe
    algorithm_options = parameters.algorithm_options
    exitwave_options = parameters.exitwave_options
    object_options = parameters.object_options
    probe_options = parameters.probe_options
    recover_probe = probe_options is not None and epoch >= probe_options.update_start
    mask = MaskInfo(exitwave_options.measured_pixels, psi.device)
    det = int(data.shape[-1])
    psi_num = None
    probe_sum = None
    batch_cost = torch.empty(algorithm_options.num_batch, dtype=torch.float32,
                             device=psi.device)
    sequence = list(range(algorithm_options.num_batch))
    stager = BatchStager(data, batches, sequence, psi.device)
    for n in sequence:
        cost, psi_num, probe_num, _ = _get_nearplane_gradients(
            stager.chunks(n), scan, psi, probe, mask, psi_num, parameters.eigen_probe,
            parameters.eigen_weights, batches, n=n, det=det,
            object_options=object_options, probe_options=probe_options,
            recover_probe=False, exitwave_options=exitwave_options, comm=comm, op=op)
        batch_cost[n] = cost
        if probe_num is not None:
            probe_sum = probe_num if probe_sum is None else probe_sum + probe_num
    stager.prefetch_next(sequence)  # DM visits the batches in the same order every epoch
    algorithm_options.costs.append([float(batch_cost.mean().item())])
    ObjectReducer(comm).finish(psi_num)
    allreduce_(comm, probe_sum)
    eps = 1e-9
    if object_options:
        # full preconditioner, no alpha mixing
        psi = psi.contiguous()
        kernels.add_quotient(psi, psi_num, object_options.preconditioner, eps)
    if recover_probe and probe_sum is not None:
        probe = probe.contiguous()
        n2 = int(probe.shape[-1] * probe.shape[-2])
        kernels.add_quotient(probe, probe_sum[0].contiguous(),
                             probe_options.preconditioner[0], eps, period=n2)
    parameters.psi = psi
    parameters.probe = probe
    return parameters
