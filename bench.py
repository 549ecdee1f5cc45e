#!/usr/bin/env python
"""Benchmark of the ptychography hot path: diffraction patterns/s per epoch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config 1..5] [--scaling weak|strong]

Default workload = BASELINE.json configs[1] (``--config 2``): rPIE, 128x128
detector, 8 probe modes, 100k scan positions per GPU, 4096x4096 complex64
object, 5 batches per epoch, synthetic data (seeded).  One "step" = one epoch
(preconditioners, all batches through the fused pipeline, object / probe
updates, multi-GPU exchange, cost read-back).  ``--config`` selects the other
BASELINE configurations (see CONFIGS); their numbers are kept under profiles/.

N > 1 is launched by torchrun (one rank per GPU, NCCL); the scan is split into
row stripes, every rank holds a replica of object and probe, and the per-batch
numerators are summed over ranks ('halo' data plane: only the object rows two
ranks share are exchanged, overlapped with the batch kernel).  ``--scaling
weak`` (default) keeps the positions per GPU fixed, ``strong`` the total.

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU oracle
(NumPy/scipy.fft restatement of the reference path, oracle/ptycho_np.py) on the
host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = 'patterns/s'
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12

# BASELINE.json `configs`, 1-based.  positions_total / gpus_nominal = positions
# per GPU of the weak-scaling run; `--scaling strong` keeps positions_total.
CONFIGS = {
    1: dict(name='lstsq_grad, 64x64 probe, 1 mode, ~1k positions, 600x600 object '
                 '(BASELINE configs[0], the CPU-runnable case)',
            algo='lstsq_grad', detector=64, modes=1, positions_total=1024, gpus_nominal=1,
            object=600, num_batch=5, cpu_sample=256),
    2: dict(name='rPIE, 128x128 detector, 8 probe modes, 100k positions per GPU, '
                 '4096x4096 complex64 object (BASELINE configs[1])',
            algo='rpie', detector=128, modes=8, positions_total=100_000, gpus_nominal=1,
            object=4096, num_batch=5, alpha=0.2, cpu_sample=192),
    3: dict(name='lstsq_grad, 256x256 detector, 4 modes, 400k positions over 8 GPUs '
                 '(50k per GPU), 4096x4096 object (BASELINE configs[2])',
            algo='lstsq_grad', detector=256, modes=4, positions_total=400_000, gpus_nominal=8,
            object=4096, num_batch=10, cpu_sample=48),
    4: dict(name='rPIE + eigen-probe variation correction + position options, 128x128 '
                 'detector, 8 modes, 200k positions (100k per GPU at 2 GPUs), 4096x4096 '
                 'object (BASELINE configs[3])',
            algo='rpie', detector=128, modes=8, positions_total=200_000, gpus_nominal=2,
            object=4096, num_batch=5, alpha=0.2, eigen=True, positions=True, cpu_sample=192),
    5: dict(name='DM, 512x512 detector (two-pass FFT), 1 mode, 16384x16384 object, 1M '
                 'positions over 8 GPUs (125k per GPU), uint16 counts '
                 '(BASELINE configs[4])',
            algo='dm', detector=512, modes=1, positions_total=1_000_000, gpus_nominal=8,
            object=16384, num_batch=25, data_dtype='uint16', cpu_sample=24),
}


def algo_bytes_per_pattern(cfg):
    """compulsory HBM bytes per pattern of the per-batch pipeline.
    rPIE / DM: N^2 s_d + 3 (N+1)^2 8 + 12; lstsq_grad adds the chi spill
    2 M N^2 8 and the phase-2 patch reads 2 (N+1)^2 8; detectors beyond shared
    memory add 4 M N^2 8 per FFT (write + read at the row / column turn)."""
    N, M = cfg['detector'], cfg['modes']
    sd = 2 if cfg.get('data_dtype') == 'uint16' else 4
    T = (N + 1) ** 2 * 8
    b = N * N * sd + 3 * T + 12
    if cfg['algo'] == 'lstsq_grad':
        b += 2 * M * N * N * 8 + 2 * T
    if N > 256:
        b += 2 * 4 * M * N * N * 8
    return b


def algo_flops_per_pattern(cfg):
    """2 M 5 N^2 log2(N^2) (FFTs) + (16 + 27 M + 40) N^2."""
    N, M = cfg['detector'], cfg['modes']
    return 2 * M * 5 * N * N * 2 * np.log2(N) + (16 + 27 * M + 40) * N * N


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', f'--id={self.index}',
                     f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits'],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE
    any pinned host buffer is allocated (first touch places the pages there),
    so every rank streams from its own memory controller and root complex."""
    import torch
    info = {'node': None}
    info['_affinity0'] = sorted(os.sched_getaffinity(0))
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f'{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
            node = int(f.read().strip())
        info['pci'] = bdf
        if node < 0:
            return info
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(node=node, cpus=len(allowed))
    except Exception as e:  # a VM without the sysfs entries: leave the affinity alone
        info['error'] = f'{type(e).__name__}: {e}'
    return info


# ------------------------------------------------------------------ data ---
def build_problem(rank, world, cfg, device):
    """Seeded synthetic experiment.  Returns host scan (global), this rank's
    split, device tensors of this rank's data, and the initial parameters."""
    import torch
    from tike_b200 import synthetic, kernels as K

    N, M, H = cfg['detector'], cfg['modes'], cfg['object']
    P = cfg['positions_per_gpu'] * world
    g = torch.Generator(device=device).manual_seed(1234)

    def smooth(shape, cutoff=0.02):
        noise = torch.randn(shape, generator=g, device=device)
        fy = torch.fft.fftfreq(shape[0], device=device)[:, None]
        fx = torch.fft.fftfreq(shape[1], device=device)[None, :]
        lp = torch.exp(-(fy * fy + fx * fx) / (2 * cutoff * cutoff))
        f = torch.fft.ifft2(torch.fft.fft2(noise) * lp).real
        del noise, lp
        f = f - f.min()
        return f / f.max()

    amp = 0.8 + 0.2 * smooth((H, H))
    phase = np.pi * (smooth((H, H)) - 0.5)
    psi_true = torch.polar(amp, phase).to(torch.complex64)[None].contiguous()
    del amp, phase
    probe = synthetic.make_probe(N, M, seed=2, photons=float(N * N) * 50.0)
    scan = synthetic.make_scan(P, H, H, N, seed=1)
    # the reference partition: equal-count row stripes, wobbly_center batches
    # (cluster.by_scan_stripes_contiguous); every rank clusters its own stripe
    from tike_b200 import cluster
    import torch.distributed as dist
    stripes = cluster.stripes_equal_count(scan, world, dim=0)
    t0 = time.perf_counter()
    part = cluster.stripe_batches(scan, stripes[rank], cfg['batch_method'], cfg['num_batch'])
    cfg['clustering_s'] = round(time.perf_counter() - t0, 1)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, part)
    else:
        parts = [part]
    order = [p[0] for p in parts]
    batches = [p[1] for p in parts]
    # what Reconstruction does with host data: neighbours back to back inside
    # every batch, positions next to another rank's stripe first (here before
    # the synthetic patterns are generated in place)
    order = cluster.band_sort_batches(scan, order, batches)
    if world > 1 and not cfg.get('positions'):
        order = cluster.boundary_first_batches(scan, order, batches, N, H)
    split = (order, batches, [p[2] for p in parts])

    local_scan = torch.as_tensor(scan[order[rank]], device=device)
    probe_d = torch.as_tensor(probe[0, 0], device=device)
    u16 = cfg.get('data_dtype') == 'uint16'
    data = torch.empty((len(local_scan), N, N),
                       dtype=torch.uint16 if u16 else torch.float32, device=device)
    step = 8192 if N <= 128 else (2048 if N <= 256 else 512)
    tmp = torch.empty((step, N, N), dtype=torch.float32, device=device) if u16 else None
    # detectors beyond shared memory run the two-pass FFT through a far-field buffer
    far = torch.empty((step, M, N, N), dtype=torch.complex64, device=device) if N > 128 else None
    for lo in range(0, len(local_scan), step):
        hi = min(len(local_scan), lo + step)
        b = K.make_batch(psi_true[0], local_scan[lo:hi].contiguous(), probe_d, N)
        fp = far[:hi - lo] if far is not None else None
        if u16:
            K.ptycho_fwd(b, fp, tmp[:hi - lo])
            data[lo:hi] = torch.clamp(torch.round(tmp[:hi - lo]), 0, 65535).to(torch.uint16)
        else:
            K.ptycho_fwd(b, fp, data[lo:hi])
    torch.cuda.synchronize()
    del psi_true, tmp, far
    psi0 = np.full((1, H, H), 0.5 + 0j, dtype=np.complex64)
    return scan, split, data, probe, psi0


def make_parameters(scan, probe, psi0, cfg, num_iter=1):
    import tike_b200.ptycho as tp
    N = cfg['detector']
    algo = cfg['algo']
    if algo == 'rpie':
        alg = tp.RpieOptions(num_batch=cfg['num_batch'], num_iter=num_iter,
                             alpha=cfg.get('alpha', 0.05), batch_method=cfg['batch_method'])
    elif algo == 'lstsq_grad':
        alg = tp.LstsqOptions(num_batch=cfg['num_batch'], num_iter=num_iter,
                              batch_method=cfg['batch_method'])
    else:
        alg = tp.DmOptions(num_batch=cfg['num_batch'], num_iter=num_iter,
                           batch_method=cfg['batch_method'])
    eigen_probe = weights = None
    if cfg.get('eigen'):
        np.random.seed(4)
        eigen_probe, weights = tp.probe.init_varying_probe(scan, probe, num_eigen_probes=2,
                                                           probes_with_modes=1)
    position_options = None
    if cfg.get('positions'):
        # rPIE ignores them exactly like the reference (rpie.py:158-170 is dead code)
        position_options = tp.PositionOptions(initial_scan=scan.copy())
    return tp.PtychoParameters(
        probe=probe.copy(), psi=psi0, scan=scan, eigen_probe=eigen_probe,
        eigen_weights=weights, algorithm_options=alg,
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((N, N), bool)),
        probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions(),
        position_options=position_options)


# --------------------------------------------------------------- cpu arm ---
def cpu_problem(cfg, sample):
    """A bounded slice of the same workload for the CPU oracle."""
    from tike_b200 import synthetic
    from oracle import ptycho_np as onp
    N, M = cfg['detector'], cfg['modes']
    H = N + 256
    psi, probe, scan = synthetic.make_problem(sample, N, M, H, H, seed=0)
    data = onp.simulate(N, probe, scan, psi)
    psi0 = np.full_like(psi, 0.5 + 0j)
    return data, scan, psi0, probe


def time_cpu_oracle(cfg, sample, steps=1, warmup=0):
    """patterns/s of the NumPy/scipy.fft port of the per-batch math (the
    gradient pass of the configured solver) on the host cores."""
    from oracle import ptycho_np as onp
    data, scan, psi, probe = cpu_problem(cfg, sample)
    mask = np.ones(data.shape[-2:], bool)
    if cfg['algo'] == 'lstsq_grad':
        pre = onp.psi_preconditioner(psi, probe, scan)

        def run():
            onp.lstsq_batch(data, scan, psi, probe, mask, pre, cfg['num_batch'])
        what = ('lstsq._get_nearplane_gradients + _precondition_nearplane_gradients '
                '(phase 1 + step lengths)')
    else:
        def run():
            onp.rpie_batch(data, scan, psi, probe, mask)
        what = 'forward + gradient math of rpie._get_nearplane_gradients'
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt, what


def metric_name(cfg):
    algo = {'rpie': 'rPIE', 'dm': 'DM'}.get(cfg['algo'], cfg['algo'])
    return (f"diffraction patterns/s per {algo} epoch "
            f"({cfg['detector']}x{cfg['detector']} detector, {cfg['modes']} probe modes)")


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cfg = dict(CONFIGS[args.config], batch_method='wobbly_center')
    sample = cfg['cpu_sample']
    value, dt, what = time_cpu_oracle(cfg, sample, steps=args.steps, warmup=min(args.warmup, 1))
    cores = os.cpu_count() or 1
    N, M = cfg['detector'], cfg['modes']
    desc = (f'{sample} positions of the same workload ({N}x{N}, {M} modes) per step, {what}; '
            'no preconditioners / updates / clustering; FFTs on all host threads '
            '(scipy.fft workers=-1), the rest NumPy')
    print(json.dumps({
        'impl': 'reference', 'metric': metric_name(cfg), 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': 'complex64', 'data': 'synthetic',
        'config': {'workload': cfg['name'] + ', CPU sample', 'config_index': args.config,
                   **{k: v for k, v in cfg.items() if k != 'name'}},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'note': 'reference CuPy path cannot run (no CuPy in image); this is the '
                'NumPy/scipy.fft port of the reference algorithm on host cores',
    }))
    return 0


# ---------------------------------------------------------------- our arm ---
def replica_checksum(t):
    """Exact integer checksum of a tensor's bits (sum of its int32 words)."""
    import torch
    words = torch.view_as_real(t).contiguous().view(torch.int32) if t.is_complex() \
        else t.contiguous().view(torch.int32)
    return int(words.to(torch.int64).sum().item())


def multi_gpu_parity(rank, world, device):
    """Small union-batch parity run carried by every multi-GPU bench line:
    `world` ranks (halo data plane) against ONE rank fed the union batches
    (mode B == single worker on concat_g(batch_g[n]))."""
    import torch
    import torch.distributed as dist
    import tike_b200.ptycho as tp
    import tike_b200.random
    from tike_b200 import synthetic
    from tike_b200.communicators import Comm
    N, M, P, H, nb, epochs = 64, 2, 1600, 400, 3, 4
    psi_t, probe, scan = synthetic.make_problem(P, N, M, H, H, seed=31)
    data = tp.simulate(N, probe, scan, psi_t)
    psi0 = np.full_like(psi_t, 0.5 + 0j)

    def params():
        return tp.PtychoParameters(
            probe=probe.copy(), psi=psi0.copy(), scan=scan.copy(),
            algorithm_options=tp.RpieOptions(num_batch=nb, num_iter=epochs, alpha=0.3),
            exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((N, N), bool)),
            probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions())

    solo_group = dist.new_group([0])  # collective: every rank calls it
    tike_b200.random.randomizer_np = np.random.default_rng(5)
    with tp.Reconstruction(data, params()) as ctx:
        order = ctx.order
        batches = ctx.comm.allgather_object([np.asarray(b).tolist() for b in ctx.batches])
        ctx.iterate(epochs)
        ctx.sync_replicas()
        sums = ctx.comm.allgather_object(
            (replica_checksum(ctx.parameters.psi), replica_checksum(ctx.parameters.probe)))
        multi = ctx.get_result()
    out = None
    if rank == 0:
        union, ranges, lo = [], [], 0
        for n in range(nb):
            idx = np.concatenate([np.asarray(order[g])[np.asarray(batches[g][n], dtype=int)]
                                  for g in range(world)])
            union.append(idx)
            ranges.append(np.arange(lo, lo + len(idx)))
            lo += len(idx)
        split = ([np.concatenate(union)], [ranges], [0])
        tike_b200.random.randomizer_np = np.random.default_rng(5)
        with tp.Reconstruction(data, params(), split=split,
                               comm=Comm(group=solo_group, single=True)) as ctx:
            ctx.iterate(epochs)
            solo = ctx.get_result()
        c_m = np.array([c[0] for c in multi.algorithm_options.costs])
        c_s = np.array([c[0] for c in solo.algorithm_options.costs])

        def rel(a, b):
            return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
        out = {'what': f'{world} ranks (halo exchange) vs 1 rank on the union batches: rPIE, '
                       f'{N}x{N}, {M} modes, {P} positions, {epochs} epochs',
               'cost_rel_err_max': float(np.max(np.abs(c_m - c_s) / np.abs(c_s))),
               'psi_rel_err': rel(multi.psi, solo.psi), 'probe_rel_err': rel(multi.probe, solo.probe),
               'tolerance': 1e-3}
        out['ok'] = bool(max(out['cost_rel_err_max'], out['psi_rel_err'],
                             out['probe_rel_err']) < 1e-3)
        out['replicas_identical'] = bool(all(s == sums[0] for s in sums))
    dist.barrier()
    return out


def time_batch_pipeline(ctx, cfg, device, repeats=4):
    """The per-batch pipeline alone (batch 0 of this rank) with CUDA events:
    ms per launch and positions per launch, for the roofline."""
    import torch
    from tike_b200 import kernels as K
    p = ctx.parameters
    N = cfg['detector']
    lo, hi = int(ctx.batches[0][0]), int(ctx.batches[0][-1]) + 1
    B = hi - lo
    ew = p.eigen_weights[lo:hi] if p.eigen_weights is not None else None
    batch = K.make_batch(p.psi[0], p.scan[lo:hi], p.probe[0, 0], N,
                         eigen_probe=p.eigen_probe[0] if p.eigen_probe is not None else None,
                         eigen_weights=ew)
    kcost = torch.empty(B, device=device)
    psi_num = torch.zeros_like(p.psi)
    probe_num = torch.empty_like(p.probe[0, 0])
    chi = None
    if cfg['algo'] == 'lstsq_grad':
        chi = torch.empty((B, 1, cfg['modes'], N, N), dtype=torch.complex64, device=device)

    def launch():
        if cfg['algo'] == 'lstsq_grad':
            K.lstsq_phase1(batch, ctx.data[lo:hi], None, N * N, noise_model='gaussian',
                           chi=chi, object_upd_sum=psi_num[0], probe_upd_sum=probe_num,
                           costs=kcost, device=device)
        else:
            K.rpie_batch(batch, ctx.data[lo:hi], None, N * N, noise_model='gaussian',
                         psi_numerator=psi_num[0], probe_numerator=probe_num,
                         costs=kcost, device=device)
    kms = []
    for it in range(repeats):
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        launch()
        k1.record()
        torch.cuda.synchronize()
        if it:
            kms.append(k0.elapsed_time(k1))
    return statistics.mean(kms), B


def run_ours(args):
    import torch
    import torch.distributed as dist
    import tike_b200.ptycho as tp
    from tike_b200 import kernels as K

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # a result-changing development switch must never leak into a measurement
    if os.environ.get('TB_DEBUG_SKIP_ALLREDUCE'):
        raise SystemExit('bench.py: TB_DEBUG_SKIP_ALLREDUCE is set; it skips the multi-GPU '
                         'reductions and would invalidate the run')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # high-priority NCCL stream: the halo exchange is enqueued while the batch
        # kernel still owns every SM and should win the first SMs that free up
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group('nccl', device_id=device, pg_options=opts)
    cfg = dict(CONFIGS[args.config], batch_method='wobbly_center')
    per_gpu = cfg['positions_total'] // cfg['gpus_nominal']
    if args.scaling == 'strong':
        per_gpu = cfg['positions_total'] // world
    if args.positions:
        per_gpu = args.positions
    cfg['positions_per_gpu'] = per_gpu
    P_total = per_gpu * world
    N, M = cfg['detector'], cfg['modes']

    parity = multi_gpu_parity(rank, world, device) if world > 1 and not args.no_parity else None

    scan, split, data, probe, psi0 = build_problem(rank, world, cfg, device)
    params = make_parameters(scan, probe, psi0, cfg)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident run ("value") ----------------------
    with tp.Reconstruction(data, params, split=split, data_is_local=True) as ctx:
        ctx.iterate(args.warmup)
        launches0 = K.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            e0.record()
            ctx.iterate(args.steps)
            ctx.sync_replicas()  # complete object replicas on every rank, inside the timing
            e1.record()
            barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = K.launch_count() - launches0
        costs = [float(np.mean(c)) for c in ctx.parameters.algorithm_options.costs]
        # every rank must hold bit-identical replicas of object and probe
        sums = (replica_checksum(ctx.parameters.psi), replica_checksum(ctx.parameters.probe))
        all_sums = ctx.comm.allgather_object(sums)
        plan = getattr(ctx.comm, 'plan', None)

        kernel_ms, B = time_batch_pipeline(ctx, cfg, device)

        # ------------ the reference's GPU op sequence with library ops -------
        standin = None
        if rank == 0 and not args.no_standin and args.config == 2:
            from baseline import torch_standin
            p = ctx.parameters
            lo = int(ctx.batches[0][0])
            nsamp = min(B, 4096)
            sl = slice(lo, lo + nsamp)
            torch_standin.rpie_batch(ctx.data[lo:lo + 256], p.scan[lo:lo + 256], p.psi[0],
                                     p.probe[0, 0])
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            torch_standin.rpie_batch(ctx.data[sl], p.scan[sl], p.psi[0], p.probe[0, 0])
            s1.record()
            torch.cuda.synchronize()
            sms = s0.elapsed_time(s1)
            standin = {'value': nsamp / (sms * 1e-3), 'unit': UNIT,
                       'sample': f'{nsamp} positions of batch 0 in 64-pattern chunks, data '
                                 'resident in HBM',
                       'what': 'PROXY, not the reference: the op sequence of '
                               'rpie._get_nearplane_gradients restated with torch.cuda library '
                               'ops (cuFFT, gather, index_add) by the authors of this repo '
                               '(baseline/torch_standin.py); the CuPy reference, its '
                               'convolution.cu and its pinned-host streaming cannot run here '
                               '(no CuPy in the image, no reference sources on the GPU box)',
                       'fused_kernel_same_sample_ratio': (B / (kernel_ms * 1e-3)) /
                                                         (nsamp / (sms * 1e-3))}

    ms_per_step = ms_total / args.steps
    value = P_total / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API, host buffers ------
    e2e = None
    want_e2e = (args.config in (1, 2, 4) or args.e2e) and not args.no_e2e
    if want_e2e:
        def run_e2e(host):
            with tp.Reconstruction(host, params, split=split, data_is_local=True,
                                   resident_data=False) as ctx:
                ctx.iterate(1)
                barrier()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                ctx.iterate(args.steps)
                ctx.sync_replicas()
                t1.record()
                barrier()
                return max_over_ranks(t0.elapsed_time(t1)) / args.steps

        # plain host -> device bandwidth of the same pinned buffer, all ranks at
        # once: the ceiling of any end-to-end number on this box
        host = torch.empty(data.shape, dtype=data.dtype, pin_memory=True)
        host.copy_(data)
        torch.cuda.synchronize()
        probe_rows = min(len(host), max(1, (1 << 30) // (N * N * host.element_size())))
        dst = torch.empty_like(data[:probe_rows])
        dst.copy_(host[:probe_rows], non_blocking=True)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(3):
            dst.copy_(host[:probe_rows], non_blocking=True)
        h1.record()
        barrier()
        h2d_ms = max_over_ranks(h0.elapsed_time(h1)) / 3
        h2d_gbs = probe_rows * N * N * host.element_size() / (h2d_ms * 1e-3) / 1e9
        del dst
        bytes_per_step = int(host.numel() * host.element_size() * world)
        f32 = None
        if host.dtype == torch.float32 and not args.e2e_float32_only:
            # The reference keeps <= 16-bit detector counts in host memory and on
            # the wire (ptycho.py:383-389) and so does this path: the headline
            # end-to-end number streams the same patterns rounded to uint16
            # counts (converted to float32 inside the kernels); the float32
            # stream is reported next to it.
            host16 = torch.empty(data.shape, dtype=torch.uint16, pin_memory=True)
            host16.copy_(torch.clamp(torch.round(data), 0, 65535).to(torch.uint16))
            torch.cuda.synchronize()
        else:
            host16 = None
        del data
        torch.cuda.empty_cache()
        e2e_ms = run_e2e(host)
        what = ('tike_b200.ptycho.Reconstruction.iterate(1) with pinned host diffraction '
                'data re-streamed every epoch (resident_data=False)')
        if host16 is not None:
            del host
            f32 = {'value': P_total / (e2e_ms * 1e-3), 'ms_per_step': e2e_ms,
                   'h2d_bytes_per_step': bytes_per_step,
                   'what': 'same run streaming the patterns as float32'}
            e2e_ms = run_e2e(host16)
            bytes_per_step //= 2
            what += ('; patterns held and streamed as uint16 counts like the reference '
                     'does for <= 16-bit data (ptycho.py:383-389)')
        e2e = {'value': P_total / (e2e_ms * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': bytes_per_step,
               'd2h_bytes_per_step': 4 * world, 'ms_per_step': e2e_ms,
               'steps': args.steps, 'what': what,
               'float32': f32,
               'h2d_probe': {'gb_per_s_per_rank': h2d_gbs, 'gb_per_s_all_ranks': h2d_gbs * world,
                             'what': 'cudaMemcpyAsync of 1 GiB of the same pinned buffer on '
                                     'every rank at once, max time over ranks: the host-side '
                                     'ceiling of e2e at this N',
                             'needed_gb_per_s_per_rank_at_value': {
                                 'float32': per_gpu * N * N * 4 / (ms_per_step * 1e-3) / 1e9,
                                 'uint16': per_gpu * N * N * 2 / (ms_per_step * 1e-3) / 1e9}},
               'numa': {k: v for k, v in numa.items() if not k.startswith('_')}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'rpie_p3_traffic.json')
    if args.config == 2 and os.path.exists(tpath):
        with open(tpath) as f:
            # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full
            # capture of this launch (20 000 positions of this workload)
            tj = json.load(f)
            traffic = tj['dram_bytes_per_pattern'] * B
    bpp = algo_bytes_per_pattern(cfg)
    fpp = algo_flops_per_pattern(cfg)
    algo_bytes = bpp * B
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    tflops = fpp * B / (kernel_ms * 1e-3) / 1e12
    cpu = None
    os.sched_setaffinity(0, numa.pop('_affinity0'))  # the CPU baseline may use every core
    if world == 1 and not args.no_cpu:
        v, dt, what = time_cpu_oracle(cfg, cfg['cpu_sample'], steps=1, warmup=0)
        cpu = {'value': v, 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
               'sample': f"{cfg['cpu_sample']} positions of the same workload, one batch of the "
                         f'NumPy/scipy.fft oracle: {what} ({dt:.1f} s)'}
    kernel_name = {'rpie': 'tb_rpie_batch', 'dm': 'tb_rpie_batch', 'lstsq_grad': 'tb_lstsq_phase1'}[
        cfg['algo']] + ((': rpie_p3_kernel (csrc/rpie_p3.cu)' if N == 128 else
                         ': rpie_fast_kernel<%d>' % N) if N <= 128 else
                        (': large-detector pipeline, register-resident K1 + K2 + K3 '
                         '(csrc/large_k13r.cu, large_k2r.cu)'
                         if N == 256 or (N == 512 and cfg['modes'] == 1) else
                         ': large-detector pipeline K1 + K2 + K3 (csrc/large_fused.cu)'))
    line = {
        'metric': metric_name(cfg), 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'complex64', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'config_index': args.config,
                   **{k: v for k, v in cfg.items() if k != 'name'},
                   'positions_total': P_total,
                   'l2': 'inputs exceed L2 (%.1f GB of patterns per GPU per epoch)' % (
                       per_gpu * N * N * (2 if cfg.get('data_dtype') == 'uint16' else 4) / 1e9),
                   'multi_gpu': ("replicated object/probe; numerators summed over ranks on the "
                                 "object rows two ranks share (halo exchange on a side stream "
                                 "under the batch kernel), probe numerator and cost by NCCL "
                                 "all-reduce") if world > 1 else 'single GPU'},
        'clocks': clocks.summary(),
        'e2e': e2e,
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                     'traffic': traffic, 'peak_kind': peak_kind,
                     'kernel': kernel_name, 'kernel_ms': kernel_ms,
                     # the fused 128^2 kernel sits above the FP32 ridge,
                     # so both components are reported; FP32 peak 148 SM x 128 lanes x 2
                     # x 1.965 GHz
                     'fp32': {'achieved_tflops': tflops, 'peak_tflops': FP32_PEAK_TFLOPS,
                              'frac': tflops / FP32_PEAK_TFLOPS},
                     'dram_frac': (traffic / (kernel_ms * 1e-3) / 1e9 / peaks['hbm_gbs'])
                     if traffic else None,
                     'algorithmic_bytes': algo_bytes,
                     'patterns_per_launch': B,
                     'algorithmic_bytes_per_pattern': bpp,
                     'algorithmic_flops_per_pattern': fpp,
                     'note': 'the fused 128^2 kernel is FP32-issue / shared-memory / L2 bound, '
                             'not HBM bound (AI ~49 FLOP/B >> ridge 11); the large-detector '
                             'pipeline (>= 256^2) is HBM bound; see DESIGN.md'},
        'cpu_baseline': cpu,
        'gpu_standin': standin,
        'cost_first_last': [costs[0], costs[-1]],
        'replicas': {'psi_probe_checksums_rank0': list(all_sums[0]),
                     'identical_across_ranks': bool(all(s == all_sums[0] for s in all_sums)),
                     'ranks': world},
        'multi_gpu_parity': parity,
        'row_plan': ({'touched': plan.touched, 'bounds': plan.bounds} if plan is not None
                     else None),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS),
                    help='BASELINE.json configuration, 1-based (default 2: the headline)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: positions per GPU fixed; strong: total positions fixed')
    ap.add_argument('--positions', type=int, default=0,
                    help='positions per GPU (default: from the configuration)')
    ap.add_argument('--e2e', action='store_true', help='force the host-streamed run')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-float32-only', action='store_true',
                    help='stream float32 patterns in the end-to-end run instead of uint16 counts')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-parity', action='store_true',
                    help='skip the small union-batch parity run of multi-GPU lines')
    ap.add_argument('--no-standin', action='store_true',
                    help='skip the torch.cuda proxy of the reference GPU op sequence')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
