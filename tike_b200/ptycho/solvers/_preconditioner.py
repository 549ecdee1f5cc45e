"""Object and probe preconditioners, recomputed once per epoch
(reference: src/tike/ptycho/solvers/_preconditioner.py:48-209).

psi preconditioner  L_O = scatter_s(sum_m |P_m|^2)   (D, H, W) complex64
probe preconditioner L_P = sum_s |patch_s(psi)|^2     (D, N, N) complex64
Each is one kernel over all positions of this worker (the reference loops
over 64-position chunks of Patch.adj / Patch.fwd)."""
from __future__ import annotations

import torch

from ... import kernels
from ._common import ObjectReducer, allreduce_


def _psi_preconditioner(parameters, streams=None, *, operator=None, order=None):
    psi = parameters.psi
    out = torch.empty_like(psi)
    if psi.shape[0] == 1:
        kernels.precond_psi(parameters.probe[0, 0], parameters.scan, out[0],
                            order=order)
        return out
    # slices >= 1 see the probe propagated through the slices before them,
    # position by position (_preconditioner.py:76-94)
    if operator is None:
        raise ValueError('the multislice object preconditioner needs the operator')
    det = int(operator.detector_shape)
    batch = kernels.multislice_batch(psi.contiguous(), parameters.scan,
                                     parameters.probe[0, 0], det, operator.norm)
    kernels.multislice_precond_psi(batch, int(psi.shape[0]),
                                   operator.fresnel_propagator(psi.device), out)
    return out


def _probe_preconditioner(parameters, streams=None, *, operator=None, order=None):
    psi = parameters.psi
    n = parameters.probe.shape[-1]
    out = torch.empty((psi.shape[0], n, n), dtype=torch.complex64, device=psi.device)
    for i in range(psi.shape[0]):  # one per slice (_preconditioner.py:131-139)
        kernels.precond_probe(psi[i], parameters.scan, out[i], order=order)
    return out


def update_preconditioners(comm, parameters, operator=None):
    """Replace (not average) both preconditioners
    (_preconditioner.py:13-37, 170-209).  ``parameters`` is one
    PtychoParameters or a list of them (one per worker, like the reference);
    with a multi-rank ``comm`` the sums run over all ranks' positions."""
    many = isinstance(parameters, (list, tuple))
    plist = list(parameters) if many else [parameters]
    for p in plist:
        both = bool(p.object_options) and bool(p.probe_options) and p.psi.is_cuda
        probe_pre = None
        # neighbouring positions back to back: the kernels keep the overlap of
        # consecutive footprints in shared memory (csrc/precond.cu)
        order = _band_order_of(p) if p.psi.is_cuda else None
        if both:
            # the two sums are independent and bound by different units (L2
            # atomics vs. gathers): run the probe one on a side stream
            current = torch.cuda.current_stream(p.psi.device)
            side = _side_stream(p.psi.device)
            side.wait_stream(current)
            order.record_stream(side)
            with torch.cuda.stream(side):
                probe_pre = _probe_preconditioner(p, operator=operator, order=order)
        if p.object_options:
            pre = _psi_preconditioner(p, operator=operator, order=order)
            reducer = ObjectReducer(comm)
            reducer.finish(pre)
            if reducer.active and reducer.plan is not None:
                # rows this rank does not touch only hold its own (zero)
                # contribution, so max(preconditioner) -- which every update
                # formula uses (rpie.py:233-236, lstsq.py:605-616) -- is a
                # maximum over ranks; the update kernels take it from here
                mx = torch.zeros(pre.shape[0], dtype=torch.float32, device=pre.device)
                for t in range(pre.shape[0]):
                    kernels.max_real(pre[t], out=mx[t:t + 1])
                comm.allreduce_max_(mx)
                pre._tb_max = mx
            p.object_options.preconditioner = pre
        if p.probe_options:
            if both:
                current.wait_stream(side)
                probe_pre.record_stream(current)
            else:
                probe_pre = _probe_preconditioner(p, operator=operator, order=order)
            allreduce_(comm, probe_pre)
            p.probe_options.preconditioner = probe_pre
    return plist if many else plist[0]


def _band_order_of(p):
    """kernels.band_order(p.scan), cached on the parameters while the scan
    tensor is unchanged (it only moves with position correction)."""
    scan = p.scan
    if p.position_options is not None:
        return kernels.band_order(scan)  # a new scan tensor every epoch
    key = (scan.data_ptr(), scan._version, tuple(scan.shape))
    cached = getattr(p, '_band_order_cache', None)
    if cached is None or cached[0] != key:
        cached = (key, kernels.band_order(scan))
        p._band_order_cache = cached
    return cached[1]


_SIDE_STREAMS = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]
