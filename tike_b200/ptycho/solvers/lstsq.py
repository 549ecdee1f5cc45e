"""Least-squares maximum-likelihood solver (Odstrcil et al. 2018)
(reference: src/tike/ptycho/solvers/lstsq.py:25-858).

Per batch: phase 1 = fused pipeline kernel that also spills the
back-propagated residual chi (csrc/rpie.cu), then the O(object) preconditioner
kernel, phase 2 = one kernel with the five per-position sums of the 2x2
step-length system (csrc/lstsq.cu).  The reference's batch-sized temporaries
``bpatches``, ``bunique_probe`` and ``bprobe_update`` (lstsq.py:394-410) are
recomputed on the fly instead of stored.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ... import kernels, linalg, opt
from ... import random as tb_random
from ..._array import to_device, to_host
from ..position import gaussian_gradient_taps
from ._common import (BatchStager, MaskInfo, ObjectReducer, allreduce_, draw_sequence, own_costs,
                      peek_sequence,
                      precond_max_of)

logger = logging.getLogger(__name__)

_ALPHA = 0.05  # fixed regularisation of lstsq.py:605-616, 635, 770


def lstsq_grad(parameters, data, batches, streams=None, worker_index=0, *,
               op, epoch, comm=None):
    """One lstsq_grad epoch; signature and side effects of lstsq.py:25-294."""
    scan, psi, probe = parameters.scan, parameters.psi, parameters.probe
    algorithm_options = parameters.algorithm_options
    eigen_weights, eigen_probe = parameters.eigen_weights, parameters.eigen_probe
    exitwave_options = parameters.exitwave_options
    position_options = parameters.position_options
    object_options = parameters.object_options
    probe_options = parameters.probe_options
    recover_probe = probe_options is not None and epoch >= probe_options.update_start
    recover_psi = object_options is not None

    # Multislice objects: the fork runs the multislice forward model and then
    # takes every gradient for slice 0 only (lstsq.py:422-530, esp. 512-530);
    # reproduced as is (csrc/multislice.cu, tb_multislice_lstsq_phase1).
    nslices = int(psi.shape[0])
    dev = psi.device
    mask = MaskInfo(exitwave_options.measured_pixels, dev)
    det = int(data.shape[-1])
    num_batch = algorithm_options.num_batch
    compact = algorithm_options.batch_method == 'compact'
    sequence, next_sequence = draw_sequence(num_batch, compact, comm)

    object_combined_update = torch.zeros_like(psi)
    probe_combined_update = torch.zeros_like(probe)
    pos_num = pos_den = None
    if position_options is not None:
        pos_num = torch.zeros_like(scan)
        pos_den = torch.zeros_like(scan)
    taps = gaussian_gradient_taps(0.333) if position_options is not None else None

    batch_cost = torch.empty(num_batch, dtype=torch.float32, device=dev)
    beta_object, beta_probe = [], []
    probe = probe.clone()
    # multi-GPU: object-sized sums are exchanged only on the rows two ranks
    # share, starting once this rank's boundary positions (before the cut of
    # the batch) are done -- see _common.ObjectReducer
    reducer = ObjectReducer(comm)
    cuts = getattr(comm, 'batch_cuts', None) if reducer.plan is not None else None
    stager = BatchStager(data, batches, sequence, dev)
    for seq_k, batch_index in enumerate(sequence):
        lo, hi = int(batches[batch_index][0]), int(batches[batch_index][-1]) + 1
        B = hi - lo
        M, N = probe.shape[-3], probe.shape[-1]
        dchunk = stager.get(seq_k)
        chi = torch.empty((B, 1, M, N, N), dtype=torch.complex64, device=dev)
        costs = torch.empty(B, dtype=torch.float32, device=dev)
        object_upd_sum = torch.zeros_like(psi) if recover_psi else None
        probe_upd_sum = torch.empty_like(probe) if recover_probe else None
        probe_part = None

        def make(plo, phi, slices=1):
            common = dict(
                eigen_probe=eigen_probe[0] if eigen_probe is not None else None,
                eigen_weights=eigen_weights[plo:phi] if eigen_weights is not None else None)
            if slices > 1:
                return kernels.multislice_batch(
                    psi.contiguous(), scan[plo:phi], probe[0, 0], det,
                    exitwave_options.propagation_normalization, **common)
            return kernels.make_batch(psi[0], scan[plo:phi], probe[0, 0], det,
                                      exitwave_options.propagation_normalization, **common)

        cut = min(max(int(cuts[batch_index]), lo), hi) if cuts is not None else hi
        pieces = [(lo, cut), (cut, hi)] if lo < cut < hi else [(lo, hi)]
        first = True
        for plo, phi in pieces:
            if phi > plo:
                target = probe_upd_sum
                if recover_probe and not first:
                    # the kernel overwrites its probe sum: pieces are added here
                    probe_part = torch.empty_like(probe) if probe_part is None else probe_part
                    target = probe_part
                kernels.lstsq_phase1(
                    make(plo, phi, nslices), dchunk[plo - lo:phi - lo], mask.dev, mask.count,
                    nslices=nslices,
                    propagator=op.fresnel_propagator(dev) if nslices > 1 else None,
                    noise_model=exitwave_options.noise_model,
                    step_mode=exitwave_options.step_length_usemodes,
                    step_length_start=exitwave_options.step_length_start,
                    step_length_weight=exitwave_options.step_length_weight,
                    unmeasured_scaling=exitwave_options.unmeasured_pixels_scaling,
                    chi=chi[plo - lo:phi - lo],
                    object_upd_sum=object_upd_sum[0] if recover_psi else None,
                    probe_upd_sum=target[0, 0] if recover_probe else None,
                    costs=costs[plo - lo:phi - lo],
                    position_num=pos_num[plo:phi] if pos_num is not None else None,
                    position_den=pos_den[plo:phi] if pos_den is not None else None,
                    taps=taps, device=dev)
                if recover_probe and not first:
                    probe_upd_sum += probe_part
                first = False
            if phi >= cut:
                reducer.begin(object_upd_sum)
        if first and recover_probe:
            probe_upd_sum.zero_()  # empty batch on this rank
        batch = make(lo, hi)

        nb_total = B
        cost_sum = costs.sum()
        if comm is not None and comm.size > 1:
            pair = torch.stack([cost_sum, torch.tensor(float(B), device=dev)])
            allreduce_(comm, probe_upd_sum, pair)
            reducer.finish(object_upd_sum)
            cost_sum, nb_total = pair[0], pair[1]
        m_probe_update = probe_upd_sum / num_batch if recover_probe else None

        # The step-length solve reads the per-position probe through the live
        # eigen_weights / eigen_probe pointers of `batch`; the reference uses
        # the snapshot `bunique_probe` taken before the eigen update
        # (lstsq.py:394-410, 654-656), so it runs first here.  It does not
        # depend on anything _update_nearplane changes.
        object_update_precond, bbeta_object, bbeta_probe = \
            _precondition_nearplane_gradients(
                batch, chi, object_upd_sum, m_probe_update,
                object_options.preconditioner if recover_psi else None,
                recover_psi=recover_psi, recover_probe=recover_probe,
                comm=comm)

        if recover_probe and eigen_weights is not None:
            eigen_probe, eigen_weights = _update_nearplane(
                chi, m_probe_update, probe, psi, scan, eigen_probe,
                eigen_weights, lo, hi, num_batch=num_batch, comm=comm)

        if recover_psi:
            if not compact:
                if object_options.use_adaptive_moment:
                    # opt.momentum (opt.py:67-82) and the step in one pass
                    psi = psi.contiguous()
                    if object_options.m is None:
                        object_options.m = torch.zeros_like(psi)
                    kernels.momentum_update(psi, object_update_precond, object_options.m,
                                            object_options.mdecay, bbeta_object.reshape(1))
                    object_options.v = None
                else:
                    psi = psi.contiguous()
                    kernels.caxpy(psi, object_update_precond, 1.0,
                                  a_dev=bbeta_object.reshape(1))
            else:
                object_combined_update += object_upd_sum
            beta_object.append(bbeta_object)

        if recover_probe:
            dprobe = bbeta_probe * m_probe_update
            probe_combined_update += dprobe / num_batch
            probe += dprobe
            beta_probe.append(bbeta_probe)

        batch_cost[batch_index] = cost_sum / nb_total

    if (position_options is not None and pos_num is not None):
        scan, position_options = _update_position(
            scan, position_options, pos_num, pos_den, epoch=epoch)

    stager.prefetch_next(peek_sequence(num_batch, compact, comm, next_sequence))
    algorithm_options.costs.append([float(batch_cost.mean().item())])

    if recover_psi and compact:
        object_update_precond = _precondition_object_update(
            object_combined_update, object_options.preconditioner,
            precond_max=precond_max_of(object_options.preconditioner))
        beta_o = torch.mean(torch.stack(beta_object))
        dpsi = beta_o * object_update_precond
        psi = psi + dpsi
        if object_options.use_adaptive_moment:
            dpsi, object_options.v, object_options.m = _momentum_checked(
                g=dpsi, v=object_options.v, m=object_options.m,
                mdecay=object_options.mdecay,
                errors=own_costs(algorithm_options.costs, worker_index),
                beta=beta_o, memory_length=3)
            weight = object_options.preconditioner
            weight = weight / (0.1 * weight.real.max() + weight)
            psi = psi + weight * dpsi

    if recover_probe and probe_options.use_adaptive_moment:
        beta_p = torch.mean(torch.stack(beta_probe))
        dprobe = probe_combined_update
        if probe_options.v is None:
            probe_options.v = torch.zeros((3, *dprobe.shape), dtype=dprobe.dtype,
                                          device=dev)
        if probe_options.m is None:
            probe_options.m = torch.zeros_like(dprobe)
        mode = 0
        d, v_new, m_new = _momentum_checked(
            g=dprobe[..., mode, :, :], v=probe_options.v[..., mode, :, :],
            m=probe_options.m[..., mode, :, :], mdecay=probe_options.mdecay,
            errors=own_costs(algorithm_options.costs, worker_index),
            beta=beta_p, memory_length=3)
        probe_options.v[..., mode, :, :] = v_new
        probe_options.m[..., mode, :, :] = m_new
        probe[..., mode, :, :] = probe[..., mode, :, :] + d

    parameters.scan = scan
    parameters.psi = psi
    parameters.probe = probe
    parameters.eigen_weights = eigen_weights
    parameters.eigen_probe = eigen_probe
    parameters.position_options = position_options
    return parameters


def _precondition_object_update(object_upd_sum, psi_update_denominator,
                                alpha: float = _ALPHA, precond_max=None):
    """object_upd / sqrt(((1-a) d)^2 + (a max d)^2)  (lstsq.py:605-616);
    the maximum is per slice (``precond_max``: the one over all ranks)."""
    out = torch.empty_like(object_upd_sum)
    upd = object_upd_sum.contiguous()
    for t in range(out.shape[0]):
        kernels.lstsq_precondition_object(
            out[t], upd[t], psi_update_denominator[t], alpha,
            precond_max=None if precond_max is None else precond_max[t:t + 1])
    return out


def _precondition_nearplane_gradients(batch, chi, object_upd_sum,
                                      m_probe_update, psi_update_denominator,
                                      *, recover_psi, recover_probe, m=0,
                                      comm=None):
    """Optimal step lengths of lstsq.py:619-718.  The five per-position sums
    come from one kernel; the 2x2 solve and the batch means are O(B) torch
    ops on the device (no host sync)."""
    B = int(batch.npos)
    N = int(batch.probe_width)
    dev = chi.device
    eps = np.float32(1e-9) / np.float32(N * N)
    object_update_precond = None
    if recover_psi:
        object_update_precond = _precondition_object_update(
            object_upd_sum, psi_update_denominator,
            precond_max=precond_max_of(psi_update_denominator))
    sums = torch.empty((B, 6), dtype=torch.float32, device=dev)
    kernels.lstsq_phase2(
        batch, chi,
        object_update_precond[0] if recover_psi else None,
        m_probe_update[0, 0, m].contiguous() if recover_probe else None,
        m, float(eps), sums)
    A1, A4, b1, b2 = sums[:, 0], sums[:, 1], sums[:, 2], sums[:, 3]
    A2 = torch.complex(sums[:, 4], sums[:, 5])

    def batch_mean(x):
        """mean over the union batch of all ranks"""
        if comm is None or comm.size == 1:
            return x.mean()
        pair = torch.stack([x.sum(), torch.tensor(float(x.numel()), device=dev)])
        allreduce_(comm, pair)
        return pair[0] / pair[1]

    if recover_psi:
        A1 = A1 + 0.5 * batch_mean(A1)
    if recover_probe:
        A4 = A4 + 0.5 * batch_mean(A4)
    x1 = x2 = None
    if recover_psi and recover_probe:
        A3 = A2.conj()
        determinant = A1 * A4 - A2 * A3
        x1 = -torch.conj(A2 * b2 - A4 * b1) / determinant
        x2 = torch.conj(A1 * b2 - A3 * b1) / determinant
        x1, x2 = x1.real, x2.real
    elif recover_psi:
        x1 = b1 / A1
    elif recover_probe:
        x2 = b2 / A4
    beta_object = beta_probe = None
    if recover_psi:
        beta_object = batch_mean(0.9 * torch.clamp(x1, min=0))
    if recover_probe:
        beta_probe = batch_mean(0.9 * torch.clamp(x2, min=0))
    return object_update_precond, beta_object, beta_probe


def _update_nearplane(chi, m_probe_update, probe, psi, scan, eigen_probe,
                      eigen_weights, lo, hi, *, num_batch, comm=None, m=0):
    """Variable-probe (OPR) updates of lstsq.py:297-364, 721-761 and
    probe.update_eigen_probe (probe.py:362-476) for the positions [lo, hi) of
    one batch.  Fused (csrc/eigen.cu): patches, per-position probe updates,
    residuals and projections are rebuilt per position inside two kernels per
    eigen probe; only (N, N) and (B,) arrays exist here.  Batch-wide means run
    over the union batch of all ranks, so the replicated eigen probes stay
    identical everywhere."""
    B = hi - lo
    dev = chi.device
    N = int(probe.shape[-1])
    chi = chi.contiguous()
    batch = kernels.make_batch(psi[0], scan[lo:hi], probe[0, 0], N)
    E1, M = int(eigen_weights.shape[-2]), int(eigen_weights.shape[-1])
    neig = 0
    if eigen_probe is not None and m < eigen_probe.shape[-3]:
        assert E1 == eigen_probe.shape[-4] + 1
        neig = int(eigen_probe.shape[-4])
        eigen_probe = eigen_probe.contiguous()

    def union(total, count=float(B)):
        """sum over this rank's positions -> sum and count over all ranks"""
        count = torch.tensor(count, dtype=torch.float32, device=dev)
        if comm is not None and comm.size > 1:
            allreduce_(comm, total, count)
        return total, count

    numden = torch.empty((B, 2), dtype=torch.float32, device=dev)
    mpu = m_probe_update[0, 0, m].contiguous()
    if neig == 0:
        kernels.lstsq_eigen_pass1(batch, chi, m, None, None, 0, None, None, 0, None, None,
                                  intensity_sums=numden)
    coefs = torch.zeros((B, max(neig, 1)), dtype=torch.complex64, device=dev)
    beta = min(0.1, 1.0 / num_batch)
    for c in range(1, neig + 1):
        w = eigen_weights[lo:hi, c, m]
        norm_weights, _ = union(torch.sum(torch.square(w)).reshape(1))
        if bool(norm_weights == 0):
            raise ValueError("eigen_probe weights cannot all be zero?")
        update = torch.zeros((N, N), dtype=torch.complex64, device=dev)
        kernels.lstsq_eigen_pass1(
            batch, chi, m, mpu, eigen_probe[0], c, coefs, eigen_weights,
            (lo * E1 + c) * M + m, (1.0 / norm_weights).to(torch.float32), update,
            intensity_sums=numden if c == 1 else None)
        update, count = union(update)
        update = update / count
        ep = eigen_probe[0, c - 1, m]
        ep = ep + beta * update / linalg.mnorm(update)
        ep = ep / linalg.mnorm(ep)
        eigen_probe[0, c - 1, m] = ep
        n = torch.empty(B, dtype=torch.float32, device=dev)
        d = torch.empty(B, dtype=torch.float32, device=dev)
        kernels.lstsq_eigen_pass2(batch, chi, m, mpu, eigen_probe[0], c,
                                  coefs if c + 1 < E1 else None, n, d)
        d_sum, count = union(torch.sum(d).reshape(1))
        eigen_weights[lo:hi, c, m] += n / (d + 0.1 * (d_sum / count))
    # _get_coefs_intensity (lstsq.py:721-736): main-probe intensity coefficient
    eigen_weights[lo:hi, 0, m] += 0.1 * numden[:, 0] / numden[:, 1]
    return eigen_probe, eigen_weights


def _update_position(scan, position_options, position_update_numerator,
                     position_update_denominator, *, alpha=_ALPHA, max_shift=1,
                     epoch=0):
    """Position step of lstsq.py:764-806 on (P, 2) arrays."""
    if epoch < position_options.update_start:
        return scan, position_options
    import scipy.stats
    num = to_host(position_update_numerator).astype(np.float32)
    den = to_host(position_update_denominator).astype(np.float32)
    step = num / ((1 - alpha) * den + alpha * max(den.max(), 1e-6))
    if position_options.update_magnitude_limit > 0:
        step = np.clip(step, -position_options.update_magnitude_limit,
                       position_options.update_magnitude_limit)
    step = step - scipy.stats.trim_mean(step, 0.05)
    if position_options.use_adaptive_moment:
        step, position_options.v, position_options.m = opt.adam(
            step, position_options.v, position_options.m,
            vdecay=position_options.vdecay, mdecay=position_options.mdecay)
    scan = scan - to_device(step.astype(np.float32), device=scan.device)
    return scan, position_options


def _momentum_checked(g, v, m, mdecay, errors, beta=1.0, memory_length=3,
                      vdecay=None):
    """Momentum that only engages while the cost decreases and recent update
    directions agree (lstsq.py:809-858)."""
    m = torch.zeros_like(g) if m is None else m
    previous_g = torch.zeros((memory_length, *g.shape), dtype=g.dtype,
                             device=g.device) if v is None else v
    previous_g = torch.roll(previous_g, shifts=-1, dims=0)
    previous_g[-1] = g / linalg.norm(g) * beta
    if (len(errors) > 2
            and max(errors[-3], errors[-2]) > min(errors[-2], errors[-1])):
        corr = linalg.inner(previous_g[:-1], previous_g[-1],
                            axis=(-2, -1)).real.flatten()
        if bool(torch.all(corr > 0)):
            friction, _ = opt.fit_line_least_squares(
                x=np.arange(len(corr) + 1),
                y=[0] + torch.log(corr).tolist())
            friction = 0.5 * max(-friction, 0)
            m = (1 - friction) * m + g
            return mdecay * m, previous_g, m
    return torch.zeros_like(g), previous_g, m / 2
