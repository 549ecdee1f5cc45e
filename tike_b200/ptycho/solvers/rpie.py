"""Regularised ptychographic iterative engine
(reference: src/tike/ptycho/solvers/rpie.py:26-567).

The per-batch work of ``_get_nearplane_gradients`` is ONE fused kernel launch
(csrc/rpie.cu) instead of ~60 CuPy launches per 64-pattern chunk; ``_update``
is two small kernels.  Reference quirks reproduced on purpose:

* the probe numerator is re-zeroed on every batch call (rpie.py:346-349), so
  in ``compact`` mode the end-of-epoch probe update only sees the last batch;
* the probe step uses ``alpha * max(preconditioner)`` only (rpie.py:269-280);
* the object gradient is divided by the number of modes (rpie.py:450);
* position correction is dead code in rPIE (rpie.py:158-170, 508-548).
"""
from __future__ import annotations

import logging

import torch

from ... import kernels, linalg, opt
from ... import random as tb_random
from ._common import (BatchStager, MaskInfo, ObjectReducer, allreduce_, draw_sequence, own_costs,
                      peek_sequence,
                      precond_max_of)
from .lstsq import _momentum_checked

logger = logging.getLogger(__name__)


def rpie(parameters, data, batches, streams=None, worker_index=0, *, op,
         epoch, comm=None, before_sync=None):
    """One rPIE epoch over this worker's batches; same signature and side
    effects as the reference solver (rpie.py:26-206) plus an optional
    ``comm`` for the multi-GPU gradient all-reduce and an
    optional ``before_sync`` callable run on the host after the last batch has
    been enqueued and before the cost is read back."""
    scan, psi, probe = parameters.scan, parameters.psi, parameters.probe
    algorithm_options = parameters.algorithm_options
    eigen_weights, eigen_probe = parameters.eigen_weights, parameters.eigen_probe
    exitwave_options = parameters.exitwave_options
    object_options = parameters.object_options
    probe_options = parameters.probe_options
    recover_probe = probe_options is not None and epoch >= probe_options.update_start

    mask = MaskInfo(exitwave_options.measured_pixels, psi.device)
    det = int(data.shape[-1])
    compact = algorithm_options.batch_method == 'compact'
    sequence, next_sequence = draw_sequence(algorithm_options.num_batch, compact, comm)

    psi_num = None
    probe_num = None
    batch_cost = torch.empty(algorithm_options.num_batch, dtype=torch.float32,
                             device=psi.device)
    # multi-GPU: the object numerator is summed over ranks only on the rows two
    # ranks share, and that exchange starts as soon as this rank's boundary
    # positions of the batch (the ones before the cut) are done
    reducer = ObjectReducer(comm)
    cuts = getattr(comm, 'batch_cuts', None) if reducer.plan is not None else None
    # with a row plan only these object rows are read or written on this rank
    rows = reducer.plan.active(comm.rank) if reducer.plan is not None else None
    stager = BatchStager(data, batches, sequence, psi.device, cuts=cuts)
    for k, n in enumerate(sequence):
        on_piece = None
        if not compact and reducer.plan is not None:
            cut = int(cuts[n]) if cuts is not None else None

            def on_piece(done_upto, numerator, cut=cut):
                if cut is None or done_upto >= cut:
                    reducer.begin(numerator)
        costs, psi_num, probe_num, eigen_weights = _get_nearplane_gradients(
            stager.chunks(k), scan, psi, probe, mask, psi_num, eigen_probe, eigen_weights,
            batches, n=int(n), det=det, object_options=object_options,
            probe_options=probe_options, recover_probe=recover_probe,
            exitwave_options=exitwave_options, comm=comm, op=op, on_piece=on_piece)
        batch_cost[n] = costs
        if not compact:
            reducer.finish(psi_num)
            allreduce_(comm, probe_num)
            psi, probe = _update(psi, probe, psi_num, probe_num,
                                 object_options, probe_options, recover_probe,
                                 algorithm_options, rows=rows)
            if rows is not None and psi_num is not None:
                psi_num[:, rows[0]:rows[1]].zero_()  # reused: only these rows were written
            else:
                psi_num = None
            probe_num = None

    # every batch is enqueued: host work that does not depend on this epoch's
    # result (Reconstruction.iterate: the affine fit of the positions) runs
    # here, under the kernels, before the cost read-back synchronises
    if before_sync is not None:
        before_sync()
    stager.prefetch_next(peek_sequence(algorithm_options.num_batch, compact, comm,
                                       next_sequence))
    algorithm_options.costs.append([float(batch_cost.mean().item())])

    if compact:
        reducer.finish(psi_num)
        allreduce_(comm, probe_num)
        psi, probe = _update(
            psi, probe, psi_num, probe_num, object_options, probe_options,
            recover_probe, algorithm_options,
            errors=own_costs(algorithm_options.costs, worker_index), rows=rows)

    if eigen_weights is not None:
        # rpie.py:209-214: weights / sqrt(mean over ALL positions of w^2)
        if comm is not None and comm.size > 1:
            sq = torch.sum(torch.square(eigen_weights), dim=-3, keepdim=True)
            count = torch.tensor(float(eigen_weights.shape[-3]), device=sq.device)
            allreduce_(comm, sq, count)
            eigen_weights = eigen_weights / torch.sqrt(sq / count)
        else:
            eigen_weights = eigen_weights / linalg.mnorm(eigen_weights, axis=-3,
                                                         keepdims=True)

    parameters.scan = scan
    parameters.psi = psi
    parameters.probe = probe
    parameters.eigen_weights = eigen_weights
    parameters.eigen_probe = eigen_probe
    return parameters


def _get_nearplane_gradients(chunks, scan, psi, probe, mask, psi_num,
                             eigen_probe, eigen_weights, batches, *, n, det,
                             object_options, probe_options, recover_probe,
                             exitwave_options, comm=None, op=None, on_piece=None):
    """Fused equivalent of rpie._get_nearplane_gradients (rpie.py:315-567).
    ``chunks`` yields ``(lo, hi, patterns)`` pieces of batch ``n`` already on
    the device (one piece for resident data, several when the patterns are
    streamed from the host).  Returns (mean batch cost as a 0-d device tensor,
    psi numerator, probe numerator (1, 1, 1, M, N, N), eigen_weights).
    ``on_piece(done_upto, psi_num)`` is called after every piece has been
    enqueued and once more at the end of the batch."""
    lo, hi = int(batches[n][0]), int(batches[n][-1]) + 1
    B = hi - lo
    dev = psi.device
    costs = torch.empty(B, dtype=torch.float32, device=dev)
    accumulate = bool(object_options)
    if accumulate and psi_num is None:
        psi_num = torch.zeros_like(psi)
    probe_num = torch.empty((psi.shape[0], *probe.shape), dtype=torch.complex64,
                            device=dev) if accumulate else None
    probe_part = None
    want_eig = recover_probe and eigen_weights is not None
    eig_step = torch.empty(B, dtype=torch.float32, device=dev) if want_eig else None
    first = True
    for clo, chi, dchunk in chunks:
        if chi <= clo:
            continue
        target = probe_num
        if accumulate and not first:
            # the kernel overwrites its probe numerator (rpie.py:346-349 re-zeroes
            # it per call); pieces of one batch are summed here
            if probe_part is None:
                probe_part = torch.empty_like(probe_num)
            target = probe_part
        nslices = int(psi.shape[0])
        common = dict(
            eigen_probe=eigen_probe[0] if eigen_probe is not None else None,
            eigen_weights=eigen_weights[clo:chi] if eigen_weights is not None else None)
        if nslices > 1:
            # multislice object: chunked slice loop (csrc/multislice.cu)
            batch = kernels.multislice_batch(
                psi.contiguous(), scan[clo:chi], probe[0, 0], det,
                exitwave_options.propagation_normalization, **common)
            kernels.multislice_rpie_batch(
                batch, nslices, op.fresnel_propagator(dev), dchunk, mask.dev, mask.count,
                noise_model=exitwave_options.noise_model,
                step_mode=exitwave_options.step_length_usemodes,
                step_length_start=exitwave_options.step_length_start,
                step_length_weight=exitwave_options.step_length_weight,
                unmeasured_scaling=exitwave_options.unmeasured_pixels_scaling,
                psi_numerator=psi_num if accumulate else None,
                probe_numerator=target[:, 0, 0] if accumulate else None,
                costs=costs[clo - lo:chi - lo],
                eigen_weight_step=eig_step[clo - lo:chi - lo] if want_eig else None,
                device=dev)
            if accumulate and not first:
                probe_num += probe_part
            first = False
            if on_piece is not None:
                on_piece(chi, psi_num)
            continue
        batch = kernels.make_batch(
            psi[0], scan[clo:chi], probe[0, 0], det,
            exitwave_options.propagation_normalization, **common)
        kernels.rpie_batch(
            batch, dchunk, mask.dev, mask.count,
            noise_model=exitwave_options.noise_model,
            step_mode=exitwave_options.step_length_usemodes,
            step_length_start=exitwave_options.step_length_start,
            step_length_weight=exitwave_options.step_length_weight,
            unmeasured_scaling=exitwave_options.unmeasured_pixels_scaling,
            psi_numerator=psi_num[0] if accumulate else None,
            probe_numerator=target[0, 0, 0] if accumulate else None,
            costs=costs[clo - lo:chi - lo],
            eigen_weight_step=eig_step[clo - lo:chi - lo] if want_eig else None,
            device=dev)
        if accumulate and not first:
            probe_num += probe_part
        first = False
        if on_piece is not None:
            on_piece(chi, psi_num)
    if on_piece is not None:
        on_piece(hi, psi_num)  # every rank starts the exchange before the cost reduction
    if want_eig:
        eigen_weights[lo:hi, 0, 0] += eig_step  # rpie.py:504-506
    if probe_num is None and recover_probe:
        # without object_options the reference still hands _update a zero
        # probe numerator (rpie.py:349): the probe step is a no-op
        probe_num = torch.zeros((psi.shape[0], *probe.shape), dtype=torch.complex64,
                                device=dev)
    cost_sum = costs.sum()
    if comm is not None and comm.size > 1:
        # cost of the union batch over all ranks; numerators are reduced by
        # the caller right before they are consumed by _update
        pair = torch.stack([cost_sum, torch.tensor(float(B), device=dev)])
        allreduce_(comm, pair)
        mean_cost = pair[0] / pair[1]
    else:
        mean_cost = cost_sum / B
    return mean_cost, psi_num, probe_num, eigen_weights


def _update(psi, probe, psi_update_numerator, probe_update_numerator,
            object_options, probe_options, recover_probe, algorithm_options,
            errors=None, rows=None):
    """rpie._update (rpie.py:217-312).  ``rows`` = (lo, hi) restricts the
    object step to those rows (multi-GPU row plan: the others are neither
    read nor written on this rank)."""
    r0, r1 = (0, psi.shape[-2]) if rows is None else rows
    alpha = algorithm_options.alpha
    if object_options:
        dpsi = psi_update_numerator
        pre = object_options.preconditioner
        pmax = precond_max_of(pre)  # per slice, over all ranks (or None: local)
        psi = psi.contiguous()
        if not object_options.use_adaptive_moment:
            for t in range(psi.shape[0]):  # max(preconditioner) is per slice
                if r1 > r0:
                    kernels.rpie_update_psi(
                        psi[t, r0:r1], dpsi[t, r0:r1], pre[t, r0:r1], alpha,
                        precond_max=None if pmax is None else pmax[t:t + 1])
        elif not errors:
            # plain step + ADAM step through the same denominator, one pass
            if object_options.v is None:
                object_options.v = torch.zeros(psi.shape, dtype=torch.float32,
                                               device=psi.device)
            if object_options.m is None:
                object_options.m = torch.zeros_like(psi)
            for t in range(psi.shape[0]):
                if r1 > r0:
                    kernels.rpie_update_psi_adam(
                        psi[t, r0:r1], dpsi[t, r0:r1], pre[t, r0:r1],
                        object_options.v[t, r0:r1], object_options.m[t, r0:r1],
                        alpha, object_options.vdecay, object_options.mdecay,
                        precond_max=None if pmax is None else pmax[t:t + 1])
        else:
            mx = (pre.real.amax(dim=(-2, -1), keepdim=True) if pmax is None
                  else pmax.reshape(-1, 1, 1))
            deno = (1 - alpha) * pre + alpha * mx
            psi = psi + dpsi / deno
            dpsi, object_options.v, object_options.m = _momentum_checked(
                g=dpsi, v=object_options.v, m=object_options.m,
                mdecay=object_options.mdecay, errors=errors,
                memory_length=3)
            psi = psi + dpsi / deno

    if recover_probe:
        dprobe = probe_update_numerator[0]
        if not probe_options.use_adaptive_moment:
            probe = probe.contiguous()
            kernels.rpie_update_probe(probe, dprobe,
                                      probe_options.preconditioner[0], alpha)
        else:
            deno = alpha * probe_options.preconditioner[0].real.amax(
                dim=(-2, -1), keepdim=True)
            probe = probe + dprobe / deno
            mode = 0  # only the main probe gets momentum (rpie.py:283-284)
            if errors:
                d, probe_options.v, probe_options.m = _momentum_checked(
                    g=dprobe[0, 0, mode], v=probe_options.v, m=probe_options.m,
                    mdecay=probe_options.mdecay, errors=errors, memory_length=3)
            else:
                d, probe_options.v, probe_options.m = opt.adam(
                    g=dprobe[0, 0, mode], v=probe_options.v, m=probe_options.m,
                    vdecay=probe_options.vdecay, mdecay=probe_options.mdecay)
            dprobe = dprobe.clone()
            dprobe[0, 0, mode] = d
            probe = probe + dprobe / deno
    return psi, probe
