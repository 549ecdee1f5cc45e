"""Ptychography driver: ``simulate``, ``reconstruct``, ``Reconstruction``
(reference: src/tike/ptycho/ptycho.py:95-1047).

Execution model: ONE PROCESS PER GPU.  Under ``torchrun`` (torch.distributed
initialised) every rank calls ``reconstruct`` with the same arguments; scan
positions are split into equal-count stripes exactly like the reference
(cluster.by_scan_stripes_contiguous), every rank keeps a replica of the object
and probe, and the per-batch gradient sums are combined with an NCCL
all-reduce (SURVEY.md §8e "mode B").  Without a process
group it is a plain single-GPU run.
"""
from __future__ import annotations

import copy
import logging
import time
import typing
import warnings

import numpy as np
import torch

from .. import cluster, kernels, precision
from .._array import pinned, to_device, to_host
from ..communicators import Comm
from ..operators import Ptycho
from . import object as tb_object
from . import probe as tb_probe
from . import solvers
from .solvers import _common as _solver_common
from .position import (AffineTransform, PositionOptions,
                       affine_position_regularization, check_allowed_positions)
from .probe import get_varying_probe

logger = logging.getLogger(__name__)

__all__ = ['reconstruct', 'simulate', 'Reconstruction', 'reconstruct_multigrid']

_FWD_CHUNK = 2048  # positions per forward-model launch in simulate / rescale


def _compute_intensity(operator, psi, scan, probe, eigen_weights=None,
                       eigen_probe=None, fly=1):
    """Detector intensity, all modes fused in one kernel per chunk
    (reference sums mode by mode: ptycho.py:95-124)."""
    P = scan.shape[-2]
    D = operator.detector_shape
    out = torch.empty((P // fly, D, D), dtype=torch.float32, device=psi.device)
    step = max(fly, (_FWD_CHUNK // fly) * fly)
    for lo in range(0, P, step):
        hi = min(P, lo + step)
        if eigen_weights is not None:
            unique = get_varying_probe(probe, eigen_probe, eigen_weights[lo:hi])
        else:
            unique = probe
        inten = operator.intensity(psi=psi, scan=scan[lo:hi], probe=unique)
        out[lo // fly:hi // fly] = inten.reshape(-1, fly, D, D).sum(dim=1)
    return out


def simulate(detector_shape, probe, scan, psi, fly=1, eigen_probe=None,
             eigen_weights=None, **kwargs):
    """Detector counts of simulated ptychography data (ptycho.py:128-179).

    probe (1, 1, SHARED, W, H) c64; scan (POSI, 2) f32; psi (D, WIDE, HIGH) or
    (WIDE, HIGH) c64.  Returns (POSI // fly, detector, detector) float32 on the
    host."""
    psi = np.asarray(psi) if not hasattr(psi, '__cuda_array_interface__') else psi
    if psi.ndim == 2:
        psi = psi[None]
    check_allowed_positions(scan, psi, probe.shape)
    with Ptycho(probe_shape=probe.shape[-1], detector_shape=int(detector_shape),
                nz=psi.shape[-2], n=psi.shape[-1], **kwargs) as operator:
        scan = to_device(scan, dtype='f32')
        psi = to_device(psi, dtype='c64')
        probe = to_device(probe, dtype='c64')
        eigen_weights = to_device(eigen_weights, dtype='f32')
        eigen_probe = to_device(eigen_probe, dtype='c64')
        data = _compute_intensity(operator, psi, scan, probe, eigen_weights,
                                  eigen_probe, fly)
        return to_host(data)


def reconstruct(data, parameters: solvers.PtychoParameters, num_gpu=1,
                use_mpi: bool = False, **kwargs) -> solvers.PtychoParameters:
    """Solve the ptychography problem (ptycho.py:182-254)."""
    with Reconstruction(data, parameters, num_gpu, use_mpi, **kwargs) as context:
        context.iterate(parameters.algorithm_options.num_iter)
        result = context.get_result()
    return result


class Reconstruction:
    """Context manager keeping the reconstruction state on the GPU between
    ``iterate`` calls (ptycho.py:265-610).

    Extra keywords (not in the reference): ``resident_data`` — True keeps the
    diffraction patterns in HBM for the whole run (default when they fit),
    False re-streams each batch from pinned host memory every epoch like the
    reference's stream_and_modify2.  ``band_sort`` — visit the members of
    every batch in band order (cluster.band_sort_batches; default True, not
    applied with position correction); the partition itself, bit-exact with
    the reference, stays available as ``cluster_order``.  ``split`` /
    ``data_is_local`` — a precomputed partition, and data already laid out in
    its order.  ``multi_gpu_mode`` — 'halo' (default: summed numerators,
    exchanged only on the object rows two ranks share and overlapped with the
    batch kernel), 'allreduce' (the same sums as an all-reduce of the whole
    object) or 'stripes' (the reference's halo-blended independent stripes).
    """

    # rPIE: run the per-epoch affine fit of the positions on the host while the
    # GPU works through the epoch (iterate); False keeps the reference's order
    early_position_fit = True

    def __init__(self, data, parameters: solvers.PtychoParameters, num_gpu=1,
                 use_mpi: bool = False, resident_data: typing.Optional[bool] = None,
                 split=None, data_is_local: bool = False,
                 multi_gpu_mode: str = 'halo', band_sort: bool = True, comm=None):
        if (np.any(np.asarray(data.shape) < 1) or data.ndim != 3
                or data.shape[-2] != data.shape[-1]):
            raise ValueError(
                f"data shape {data.shape} is incorrect. "
                "It should be (N, W, H), "
                "where N >= 1 is the number of square diffraction patterns.")
        if data.shape[0] != parameters.scan.shape[0] and not data_is_local:
            raise ValueError(
                f"data shape {data.shape} and scan shape {parameters.scan.shape} "
                "are incompatible. They should have the same leading dimension.")
        if np.any(np.asarray(parameters.probe.shape[-2:]) > np.asarray(data.shape[-2:])):
            raise ValueError(f"probe shape {parameters.probe.shape} "
                             f"and data shape {data.shape} are incompatible. "
                             "The probe width/height must be "
                             f"<= the data width/height .")
        logger.info("%s on %d - %d by %d frames for at most %d epochs.",
                    parameters.algorithm_options.name, *data.shape[-3:],
                    parameters.algorithm_options.num_iter)
        if isinstance(num_gpu, tuple):
            torch.cuda.set_device(num_gpu[0])
        requested = len(num_gpu) if isinstance(num_gpu, tuple) else int(num_gpu)
        world = (torch.distributed.get_world_size()
                 if torch.distributed.is_available() and torch.distributed.is_initialized() else 1)
        if comm is not None:
            world = comm.size
        if requested > 1 and world == 1:
            # the reference drives several GPUs from threads of one process
            # (pool.py:397-413); here every GPU is its own process
            warnings.warn(
                f"num_gpu={num_gpu} was requested, but this process is not part of a "
                "torch.distributed job: one GPU is used.  Launch one process per GPU, e.g. "
                f"`torchrun --nproc-per-node {requested} script.py`.", UserWarning)
        if multi_gpu_mode not in ('halo', 'allreduce', 'stripes'):
            raise ValueError("multi_gpu_mode must be 'halo', 'allreduce' or 'stripes', "
                             f"not {multi_gpu_mode!r}")
        # 'allreduce': object/probe replicated, gradient sums all-reduced per batch
        #   (equivalent to one worker seeing the union batches).
        # 'halo' (default): the same sums, but object-sized arrays are only
        #   exchanged on the rows two ranks' footprints share (about one probe
        #   height per neighbour, communicators.RowPlan) while the rank's interior
        #   positions are still being processed; every rank owns a row range of
        #   the object and replicas are refreshed once per epoch.  Same results
        #   as 'allreduce' up to float summation order.
        # 'stripes': the reference's scheme (ptycho.py:474-502) -- every rank
        #   reconstructs its stripe independently, then the probes are averaged,
        #   the object halos blended (pool.py:415-476) and the stripes stitched
        #   in get_result (object.py:154-167).
        self.multi_gpu_mode = multi_gpu_mode
        # visit neighbouring positions back to back inside every batch
        # (cluster.band_sort_batches); False keeps the clustering's own order
        self.band_sort = band_sort
        self._data_in = data
        self._parameters_in = copy.deepcopy(parameters)
        self.resident_data = resident_data
        # optional precomputed (order, batches, stripe_start) replacing the
        # host-side clustering (same structure as cluster.by_scan_stripes_contiguous)
        self._split = split
        # data_is_local: `data` already holds only this rank's patterns, in
        # the order of split[0][rank] (large multi-GPU runs cannot afford a
        # full copy of the data on every rank)
        self._data_is_local = data_is_local
        popt = parameters.probe_options
        oopt = parameters.object_options
        self.operator = Ptycho(
            probe_shape=parameters.probe.shape[-1],
            detector_shape=data.shape[-1],
            nz=parameters.psi.shape[-2], n=parameters.psi.shape[-1],
            norm=parameters.exitwave_options.propagation_normalization,
            probe_wavelength=popt.probe_wavelength if popt else float('nan'),
            probe_FOV_lengths=popt.probe_FOV_lengths if popt else (float('nan'),) * 2,
            multislice_propagation_distance=oopt.multislice_propagation_distance
            if oopt else 1e-9,
        )
        self.comm = Comm() if comm is None else comm
        self.parameters: solvers.PtychoParameters = None
        self.data = None

    # ------------------------------------------------------------------
    def __enter__(self):
        self.operator.__enter__()
        self.comm.__enter__()
        data, params = self._data_in, self._parameters_in
        host_data = None if isinstance(data, torch.Tensor) and data.is_cuda else data
        if host_data is not None:
            sample = np.asarray(host_data[:min(len(host_data), 64)])
            if not np.all(np.isfinite(sample)) or np.any(sample < 0):
                warnings.warn(
                    "Diffraction patterns contain invalid data. "
                    "All data should be non-negative and finite.", UserWarning)

        alg = params.algorithm_options
        scan_host = np.asarray(to_host(params.scan))
        if self._split is not None:
            split = self._split
        elif self.comm.size > 1 and alg.batch_method == 'wobbly_center':
            # deterministic method: every rank clusters only its own stripe
            owner = cluster.stripes_equal_count(scan_host, self.comm.size, dim=0)
            part = cluster.stripe_batches(scan_host, owner[self.comm.rank],
                                          alg.batch_method, alg.num_batch)
            parts = self.comm.allgather_object(part)
            split = ([p[0] for p in parts], [p[1] for p in parts],
                     [p[2] for p in parts])
        else:
            # methods drawing from the global NumPy generator are evaluated on
            # rank 0 in stripe order (like the reference) and broadcast
            split = None
            if self.comm.rank == 0:
                split = cluster.by_scan_stripes_contiguous(
                    scan=scan_host, num_workers=self.comm.size,
                    batch_method=alg.batch_method, num_batch=alg.num_batch)
            split = self.comm.bcast_object(split)
        self.order, batches, self.stripe_start = split
        # the partition exactly as the reference computes it; `order` below may
        # visit the members of a batch in another sequence
        self.cluster_order = self.order
        if (self.band_sort and not self._data_is_local
                and params.position_options is None):
            # neighbouring positions back to back inside every batch (same
            # batches, same ranges; every rank computes the same permutation).
            # Not with position correction: the RANSAC affine fit draws its
            # subsets by array index and its float32 normal equations are
            # sensitive to the summation order (position.py:277-327), so the
            # reference's own sequence is kept there.
            self.order = cluster.band_sort_batches(scan_host, self.order, batches)
        # object rows are only exchanged where stripes overlap ('halo'); the
        # checked momentum of compact + adaptive moments takes norms over the
        # whole object (lstsq.py:809-858), so it keeps full replicas
        oopt = params.object_options
        self._halo = (self.multi_gpu_mode == 'halo' and self.comm.size > 1 and not (
            oopt is not None and oopt.use_adaptive_moment
            and alg.batch_method == 'compact'))
        if (self._halo and not self._data_is_local
                and params.position_options is None):
            self.order = cluster.boundary_first_batches(
                scan_host, self.order, batches, params.probe.shape[-1],
                params.psi.shape[-2])
        mine = self.order[self.comm.rank]
        self.batches = batches[self.comm.rank]

        # data -> pinned host (dtype kept when <= 16 bit, ptycho.py:383-390)
        dev = torch.device('cuda', torch.cuda.current_device())
        if self._data_is_local:
            if len(data) != len(mine):
                raise ValueError('local data must match this rank\'s positions')
            if isinstance(data, torch.Tensor):
                self.data = data
            else:
                local = np.asarray(data)
                resident = self.resident_data
                if resident is None:
                    free, _ = torch.cuda.mem_get_info()
                    resident = local.nbytes < 0.6 * free
                self.data = (torch.from_numpy(np.ascontiguousarray(local)).to(dev)
                             if resident else pinned(local))
        elif isinstance(data, torch.Tensor) and data.is_cuda:
            self.data = data[torch.as_tensor(mine, device=data.device)].contiguous()
        else:
            if isinstance(data, torch.Tensor):
                data = data.numpy()
            # counts stay 16 bit in memory and on the wire like the reference
            # (ptycho.py:383-390) -- but only unsigned integers: float16 or
            # signed data would lose values in a uint16 cast
            if data.dtype == np.uint16:
                local = np.asarray(data[mine])
            elif data.dtype == np.uint8:
                local = np.asarray(data[mine]).astype(np.uint16)
            else:
                local = np.asarray(data[mine], dtype=precision.floating)
            resident = self.resident_data
            if resident is None:
                free, _ = torch.cuda.mem_get_info()
                resident = local.nbytes < 0.6 * free
            self.data = (torch.from_numpy(np.ascontiguousarray(local)).to(dev)
                         if resident else pinned(local))

        host_params = solvers.PtychoParameters.split(mine, x=params.copy_to_host())
        self.parameters = host_params.copy_to_device()

        popt = self.parameters.probe_options
        if popt is not None and popt.init_rescale_from_measurements:
            self.parameters = _rescale_probe(self.operator, self.comm, self.data,
                                             self.parameters)
        self.comm.plan = None
        self.comm.batch_cuts = None
        self._replicas_stale = False
        if self._halo:
            self._refresh_row_plan()
        return self

    def _refresh_row_plan(self):
        """Row ranges every rank touches / owns for the current scan, and per
        batch the index after this rank's last boundary position."""
        from ..communicators import RowPlan
        p = self.parameters
        rows = np.asarray(to_host(p.scan))[:, 0]
        mine = (float(rows.min()), float(rows.max())) if len(rows) else None
        width = int(p.probe.shape[-1])
        plan = RowPlan.from_scan_rows(self.comm.allgather_object(mine), width,
                                      int(p.psi.shape[-2]))
        near = cluster.boundary_mask(rows, plan.shared_rows(self.comm.rank), width)
        cuts = []
        for b in self.batches:
            if len(b) == 0:
                cuts.append(0)
                continue
            hit = np.nonzero(near[np.asarray(b)])[0]
            cuts.append(int(b[0]) + (int(hit[-1]) + 1 if len(hit) else 0))
        self.comm.plan = plan
        self.comm.batch_cuts = cuts

    # ------------------------------------------------------------------
    def iterate(self, num_iter: int) -> None:
        """Advance the reconstruction by num_iter epochs (ptycho.py:431-564)."""
        p = self.parameters
        alg = p.algorithm_options
        start = time.perf_counter()
        solver = getattr(solvers, alg.name)
        for _ in range(num_iter):
            if np.sum(alg.times) > alg.time_limit:
                logger.info("Maximum reconstruction time exceeded.")
                break
            epoch = len(alg.times)
            logger.info("%s epoch %d", alg.name, epoch)

            p = _apply_probe_constraints(p, epoch=epoch)
            if self._halo and p.position_options is not None:
                self._refresh_row_plan()  # the positions moved
            stripes = self.multi_gpu_mode == 'stripes' and self.comm.size > 1
            solver_comm = None if stripes else self.comm
            # rPIE never moves the positions (its position correction is dead
            # code in the reference), so the per-epoch affine fit of
            # the positions (ptycho.py:854-866) -- host work, ~15 ms at 100 k
            # positions -- can run while the GPU works through the epoch instead
            # of after the cost read-back: same inputs, same generator draws in
            # the same order, same result.  One process only (several ranks
            # average the transform between the solver and the fit).
            early_fit = (self.early_position_fit and alg.name == 'rpie'
                         and bool(p.position_options) and self.comm.size == 1)
            fitted = {}
            extra = {}
            if early_fit:
                scan_host = to_host(p.scan)  # nothing of this epoch is enqueued yet

                def fit(scan_host=scan_host, options=p.position_options):
                    fitted['scan'], fitted['options'] = affine_position_regularization(
                        updated=scan_host, position_options=options)
                extra['before_sync'] = fit
            p = solvers.update_preconditioners(solver_comm, p, self.operator)
            # checked momentum compares this worker's own cost history
            # (lstsq.py:255-262); alg.costs rows hold one cost per rank
            worker = self.comm.rank if stripes else 0
            p = solver(p, self.data, self.batches, None, worker, op=self.operator,
                       epoch=epoch, comm=solver_comm, **extra)
            if stripes:
                p = self._exchange_stripes(p)
            if self._halo:
                # Every rank advanced the rows under its own footprints.  The
                # next epoch only reads those rows again, so the owners hand
                # their neighbours the halo rows; complete replicas are only
                # rebuilt when something looks at the whole object -- a
                # constraint with a spatial footprint now, or the caller later
                # (sync_replicas; the reference, too, keeps one object per GPU
                # and joins them in get_result, object.py:154-167).
                p.psi = p.psi.contiguous()
                if self._needs_whole_object(p):
                    self.comm.gather_owned_rows_(p.psi, self.comm.plan)
                    self._replicas_stale = False
                else:
                    self.comm.halo_refresh_(p.psi, self.comm.plan)
                    self._replicas_stale = True

            if p.position_options is not None and self.comm.size > 1:
                buffers = self.comm.allgather_object(
                    p.position_options.transform.asbuffer())
                p.position_options.transform = AffineTransform.frombuffer(
                    np.mean(buffers, axis=0))

            p = _apply_object_constraints(p, comm=self.comm)
            if early_fit:
                p.position_options = fitted['options']
                if p.position_options.use_position_regularization:
                    p.scan = to_device(fitted['scan'], device=p.scan.device)
            else:
                p = _apply_position_constraints(p)

            # one cost per worker, like the reference (ptycho.py:531-537)
            if stripes:
                alg.costs[-1] = [c for part in self.comm.allgather_object(alg.costs[-1])
                                 for c in part][:max(1, self.comm.size)]
            else:
                # the solvers already reduced the cost over the union batches
                alg.costs[-1] = list(alg.costs[-1]) * max(1, self.comm.size)
            alg.times.append(time.perf_counter() - start)
            start = time.perf_counter()
            logger.info("%10s cost is %+1.3e", p.exitwave_options.noise_model,
                        np.mean(alg.costs[-1]))
        self.parameters = p

    @staticmethod
    def _needs_whole_object(p) -> bool:
        oopt = p.object_options
        return oopt is not None and bool(
            oopt.positivity_constraint or oopt.smoothness_constraint or oopt.clip_magnitude)

    def sync_replicas(self) -> None:
        """'halo' data plane: make every rank's object a complete, current copy
        (each row taken from its owner).  Called by get_result / get_psi /
        __exit__; a collective -- every rank must call it."""
        if getattr(self, '_replicas_stale', False) and self.parameters is not None:
            self.parameters.psi = self.comm.gather_owned_rows_(
                self.parameters.psi.contiguous(), self.comm.plan)
            self._replicas_stale = False

    def _exchange_stripes(self, p):
        """End-of-epoch exchange of the reference's multi-GPU scheme
        (ptycho.py:474-502).  The probe mean only reaches worker 0 there
        (enumerate() over a length-1 result) -- reproduced."""
        mean = p.probe.clone()
        self.comm.allreduce_mean_(mean)
        if self.comm.rank == 0:
            p.probe = mean
        if p.eigen_probe is not None:
            mean = p.eigen_probe.clone()
            self.comm.allreduce_mean_(mean)
            if self.comm.rank == 0:
                p.eigen_probe = mean
        pw = p.probe.shape[-2]
        p.psi = self.comm.swap_edges(p.psi.contiguous(), overlap=pw - 1,
                                     edges=self.stripe_start)
        return p

    # ------------------------------------------------------------------
    def _reorder(self):
        return np.argsort(np.concatenate(self.order))

    def get_scan(self):
        parts = self.comm.allgather_object(to_host(self.parameters.scan))
        return np.concatenate(parts, axis=0)[self._reorder()]

    def get_result(self) -> solvers.PtychoParameters:
        """Current estimates as host arrays in the caller's original position
        order (ptycho.py:573-597).  Object and probe are replicated over
        ranks, so they are taken from this rank."""
        self.sync_replicas()
        local = self.parameters.copy_to_host()
        reorder = self._reorder()
        scan = np.concatenate(self.comm.allgather_object(local.scan), axis=0)[reorder]
        weights = None
        if local.eigen_weights is not None:
            weights = np.concatenate(
                self.comm.allgather_object(local.eigen_weights), axis=0)[reorder]
        pos = None
        if local.position_options is not None:
            pos = PositionOptions.join(
                self.comm.allgather_object(local.position_options), reorder)
        if self.multi_gpu_mode == 'stripes' and self.comm.size > 1:
            # join_psi / x[0].probe of PtychoParameters.join (options.py:293-330)
            from ..communicators.comm import stitch_stripes
            local.psi = stitch_stripes(self.comm.allgather_object(local.psi),
                                       self.stripe_start, local.probe.shape[-2])
            local.probe, local.eigen_probe = self.comm.bcast_object(
                (local.probe, local.eigen_probe))
        return solvers.PtychoParameters(
            probe=local.probe, psi=local.psi, scan=scan,
            eigen_probe=local.eigen_probe, eigen_weights=weights,
            algorithm_options=local.algorithm_options,
            exitwave_options=local.exitwave_options,
            probe_options=local.probe_options,
            object_options=local.object_options, position_options=pos)

    def get_convergence(self):
        alg = self.parameters.algorithm_options
        return alg.costs, alg.times

    def get_psi(self):
        self.sync_replicas()
        return to_host(self.parameters.psi)

    def get_probe(self):
        p = self.parameters
        weights = None
        if p.eigen_weights is not None:
            weights = np.concatenate(
                self.comm.allgather_object(to_host(p.eigen_weights)),
                axis=0)[self._reorder()]
        return to_host(p.probe), to_host(p.eigen_probe), weights

    def append_new_data(self, new_data, new_scan) -> None:
        raise NotImplementedError(
            "Adding data on-the-fly is disabled until further notice.")

    def __exit__(self, type, value, traceback):
        if self.parameters is not None:
            if type is None:
                self.sync_replicas()
            self.parameters = self.parameters.copy_to_host()
        self.data = None
        _solver_common.release_host_rings()
        self.comm.__exit__(type, value, traceback)
        self.operator.__exit__(type, value, traceback)
        kernels.free_scratch()
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------
def _apply_probe_constraints(parameters, *, epoch: int):
    """Per-epoch probe constraints (ptycho.py:723-808)."""
    popt = parameters.probe_options
    if popt is None:
        return parameters
    if popt.recover_probe(epoch):
        if popt.probe_support > 0:
            b0 = tb_probe.finite_probe_support(
                parameters.probe, p=popt.probe_support,
                radius=popt.probe_support_radius,
                degree=popt.probe_support_degree)
            parameters.probe = parameters.probe - b0 * torch.conj(b0 * parameters.probe)
        if popt.additional_probe_penalty > 0:
            b1 = popt.additional_probe_penalty * torch.linspace(
                0, 1, parameters.probe.shape[-3], dtype=torch.float32,
                device=parameters.probe.device)[..., None, None]
            parameters.probe = parameters.probe - b1 * torch.conj(b1 * parameters.probe)
        if popt.median_filter_abs_probe:
            parameters.probe = tb_probe.apply_median_filter_abs_probe(
                parameters.probe, med_filt_px=popt.median_filter_abs_probe_px)
        if popt.force_centered_intensity:
            parameters.probe = tb_probe.constrain_center_peak(parameters.probe)
        if popt.force_sparsity < 1:
            parameters.probe = tb_probe.constrain_probe_sparsity(
                parameters.probe, f=popt.force_sparsity)
        if popt.force_orthogonality:
            parameters.probe, power = tb_probe.orthogonalize_eig(parameters.probe)
        else:
            power = tb_probe.power(parameters.probe)
        popt.power.append(power)  # device array; ProbeOptions.copy_to_host converts

    alg = parameters.algorithm_options
    if alg.rescale_method == "constant_probe_photons" and (
            len(alg.costs) % alg.rescale_period == 0):
        parameters.probe = tb_probe.rescale_probe_using_fixed_intensity_photons(
            parameters.probe, Nphotons=popt.probe_photons,
            probe_power_fraction=None)

    if parameters.eigen_probe is not None and popt.recover_probe(epoch):
        parameters.eigen_probe, parameters.eigen_weights = \
            tb_probe.constrain_variable_probe(parameters.eigen_probe,
                                              parameters.eigen_weights)
    return parameters


def _apply_object_constraints(parameters, comm=None):
    """Per-epoch object constraints (ptycho.py:811-851) as single passes over
    the object (csrc/update.cu)."""
    oopt = parameters.object_options
    if oopt is None:
        return parameters
    psi = parameters.psi
    on_device = isinstance(psi, torch.Tensor) and psi.is_cuda
    if not on_device:
        return _apply_object_constraints_host(parameters)
    if oopt.positivity_constraint and oopt.positivity_constraint > 1:
        raise ValueError("Positivity constraint must be in the range [0, 1] not "
                         f"{oopt.positivity_constraint}.")
    positivity = float(oopt.positivity_constraint or 0.0)
    if oopt.smoothness_constraint:
        if not (0 <= oopt.smoothness_constraint < 1.0 / 8.0):
            raise ValueError("Smoothness constraint must be in range [0, 1/8) not "
                             f"{oopt.smoothness_constraint}.")
        psi = psi.contiguous()
        if positivity > 0:
            kernels.object_pointwise_constraints(psi, positivity=positivity)
        psi = kernels.object_smoothness(psi, oopt.smoothness_constraint)
        if oopt.clip_magnitude:
            kernels.object_pointwise_constraints(psi, clip=True, a_max=1.0)
    elif positivity > 0 or oopt.clip_magnitude:
        psi = psi.contiguous()
        kernels.object_pointwise_constraints(psi, positivity=positivity,
                                             clip=bool(oopt.clip_magnitude), a_max=1.0)
    parameters.psi = psi
    alg = parameters.algorithm_options
    if (alg.name != "dm" and alg.rescale_method == "mean_of_abs_object"
            and oopt.preconditioner is not None
            and len(alg.costs) % alg.rescale_period == 0):
        parameters.psi, parameters.probe = _remove_object_ambiguity(
            parameters.psi.contiguous(), parameters.probe.contiguous(),
            oopt.preconditioner, comm)
    return parameters


def _remove_object_ambiguity(psi, probe, preconditioner, comm=None):
    """object.py:324-335 with the two reductions in one kernel and the
    scaling read from the device (no host synchronisation).  With the object
    rows split over ranks every rank sums its own rows."""
    plan = getattr(comm, 'plan', None) if comm is not None and comm.size > 1 else None
    if plan is None:
        sums = kernels.weighted_norm_sums(psi, preconditioner)
    else:
        lo, hi = plan.own(comm.rank)
        sums = torch.zeros(2, dtype=torch.float64, device=psi.device)
        if hi > lo:
            for t in range(psi.shape[0]):
                sums += kernels.weighted_norm_sums(psi[t, lo:hi], preconditioner[t, lo:hi])
        comm.allreduce_sum_(sums)
    n = float(psi.numel())
    # W / mnorm(W); 2 sqrt(mean(|psi|^2 W))
    object_norm = (2.0 * torch.sqrt((sums[0] / n) / torch.sqrt(sums[1] / n))).to(
        torch.float32).reshape(1)
    kernels.scale_by_device_scalar(psi, object_norm, divide=True)
    kernels.scale_by_device_scalar(probe, object_norm, divide=False)
    return psi, probe


def _apply_object_constraints_host(parameters):
    """The same constraints through the array expressions of ptycho/object.py
    (host arrays; used by callers that apply them outside a reconstruction)."""
    oopt = parameters.object_options
    if oopt.positivity_constraint:
        parameters.psi = tb_object.positivity_constraint(
            parameters.psi, r=oopt.positivity_constraint)
    if oopt.smoothness_constraint:
        parameters.psi = tb_object.smoothness_constraint(
            parameters.psi, a=oopt.smoothness_constraint)
    if oopt.clip_magnitude:
        parameters.psi = tb_object.clip_magnitude(parameters.psi, a_max=1.0)
    alg = parameters.algorithm_options
    if (alg.name != "dm" and alg.rescale_method == "mean_of_abs_object"
            and oopt.preconditioner is not None
            and len(alg.costs) % alg.rescale_period == 0):
        parameters.psi, parameters.probe = tb_object.remove_object_ambiguity(
            parameters.psi, parameters.probe, oopt.preconditioner)
    return parameters


def _apply_position_constraints(parameters):
    """Affine regularisation of positions, every epoch (ptycho.py:854-866)."""
    if parameters.position_options:
        parameters.scan, parameters.position_options = \
            affine_position_regularization(
                updated=parameters.scan,
                position_options=parameters.position_options)
    return parameters


def _get_rescale(data, parameters, operator):
    """Sum of measured and of modelled intensity over measured pixels in
    float64 (ptycho.py:873-918)."""
    from .solvers._common import MaskInfo, stage_data
    dev = parameters.psi.device
    mask = MaskInfo(parameters.exitwave_options.measured_pixels, dev)
    m = None if mask.all else mask.dev.bool()
    sums = torch.zeros(2, dtype=torch.float64, device=dev)
    P = parameters.scan.shape[0]
    for lo in range(0, P, _FWD_CHUNK):
        hi = min(P, lo + _FWD_CHUNK)
        inten = operator.intensity(psi=parameters.psi,
                                   scan=parameters.scan[lo:hi],
                                   probe=parameters.probe)
        d = stage_data(data, lo, hi, dev).to(torch.float32)
        if m is None:
            sums[0] += d.sum(dtype=torch.float64)
            sums[1] += inten.sum(dtype=torch.float64)
        else:
            sums[0] += d[:, m].sum(dtype=torch.float64)
            sums[1] += inten[:, m].sum(dtype=torch.float64)
    return sums.cpu().numpy()


def _rescale_probe(operator, comm, data, parameters):
    """Scale the probe so modelled and measured intensity sums match
    (ptycho.py:921-972)."""
    try:
        n = _get_rescale(data, parameters, operator)
    except torch.cuda.OutOfMemoryError:
        raise ValueError(
            "tike.ptycho.reconstruct ran out of memory! "
            "Increase num_batch to process your data in smaller chunks.")
    n = np.sqrt(comm.reduce_cpu_sum(n))
    rescale = np.float32(n[0] / n[1])
    logger.info("Probe rescaled by %f", rescale)
    parameters.probe = parameters.probe * float(rescale)
    popt = parameters.probe_options
    if np.isnan(popt.probe_photons):
        popt.probe_photons = float(torch.sum(torch.square(parameters.probe.abs())).item())
    return parameters


def reconstruct_multigrid(data, parameters, num_gpu=1, use_mpi=False,
                          num_levels: int = 3, interp=None):
    """Multi-grid reconstruction: coarse-to-fine with Fourier-cropped data
    (ptycho.py:975-1047)."""
    interp = solvers.options._resize_fft if interp is None else interp
    if (data.shape[-1] * 0.5**(num_levels - 1)) < 64:
        warnings.warn('Cropping diffraction patterns to less than 64 pixels '
                      'wide is not recommended because the full doughnut'
                      ' may be visible.')
    resampled = parameters.resample(0.5**(num_levels - 1), interp)
    for level in range(num_levels - 1, -1, -1):
        level_data = data if level == 0 else solvers.crop_fourier_space(
            data, data.shape[-1] // (2**level))
        with Reconstruction(data=level_data, parameters=resampled,
                            num_gpu=num_gpu, use_mpi=use_mpi) as context:
            context.iterate(resampled.algorithm_options.num_iter)
            result = context.get_result()
        if level == 0:
            return result
        resampled = result.resample(2.0, interp)
    raise RuntimeError('This should not happen.')
