#!/usr/bin/env python
"""Headline benchmark: diffraction patterns/s per rPIE epoch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): rPIE, 128x128 detector, 8 probe modes,
100k scan positions per GPU, 4096x4096 complex64 object, 5 batches per epoch,
synthetic data (seeded).  One "step" = one rPIE epoch (preconditioners, all
batches through the fused kernel, object/probe updates, cost read-back).

N > 1 is launched by torchrun (one rank per GPU, NCCL); the scan is split into
row stripes, every rank holds a replica of object and probe and the per-batch
numerators are all-reduced ("scaling": "weak": 100k positions per GPU).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle
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

METRIC = 'diffraction patterns/s per rPIE epoch (128x128 detector, 8 probe modes)'
UNIT = 'patterns/s'
WORKLOAD = dict(detector=128, modes=8, positions_per_gpu=100_000, object=4096,
                num_batch=5, alpha=0.2, batch_method='wobbly_center')
# compulsory HBM bytes per pattern of the fused rPIE batch
# kernel at N = 128, float32 data: N^2*4 + 3*(N+1)^2*8 + 12
ALGO_BYTES_PER_PATTERN = 128 * 128 * 4 + 3 * 129 * 129 * 8 + 12
# 2 * M * 5 N^2 log2(N^2) (FFTs) + (16 + 27 M + 40) N^2 (elementwise)
ALGO_FLOPS_PER_PATTERN = 2 * 8 * 5 * 128 * 128 * 14 + (16 + 27 * 8 + 40) * 128 * 128
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


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
        f = f - f.min()
        return f / f.max()

    amp = 0.8 + 0.2 * smooth((H, H))
    phase = np.pi * (smooth((H, H)) - 0.5)
    psi_true = torch.polar(amp, phase).to(torch.complex64)[None].contiguous()
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
    # what Reconstruction does with host data: neighbours back to back inside
    # every batch (here before the synthetic patterns are generated in place)
    order = cluster.band_sort_batches(scan, order, [p[1] for p in parts])
    split = (order, [p[1] for p in parts], [p[2] for p in parts])

    local_scan = torch.as_tensor(scan[order[rank]], device=device)
    probe_d = torch.as_tensor(probe[0, 0], device=device)
    data = torch.empty((len(local_scan), N, N), dtype=torch.float32, device=device)
    for lo in range(0, len(local_scan), 8192):
        hi = min(len(local_scan), lo + 8192)
        b = K.make_batch(psi_true[0], local_scan[lo:hi].contiguous(), probe_d, N)
        K.ptycho_fwd(b, None, data[lo:hi])
    torch.cuda.synchronize()
    psi0 = np.full((1, H, H), 0.5 + 0j, dtype=np.complex64)
    return scan, split, data, probe, psi0


def make_parameters(scan, probe, psi0, cfg):
    import tike_b200.ptycho as tp
    N = cfg['detector']
    return tp.PtychoParameters(
        probe=probe.copy(), psi=psi0, scan=scan,
        algorithm_options=tp.RpieOptions(num_batch=cfg['num_batch'], num_iter=1,
                                         alpha=cfg['alpha'],
                                         batch_method=cfg['batch_method']),
        exitwave_options=tp.ExitWaveOptions(measured_pixels=np.ones((N, N), bool)),
        probe_options=tp.ProbeOptions(), object_options=tp.ObjectOptions())


# --------------------------------------------------------------- cpu arm ---
def cpu_sample(cfg, sample):
    """(inputs) a bounded slice of the same workload for the CPU oracle."""
    from tike_b200 import synthetic
    from oracle import ptycho_np as onp
    N, M = cfg['detector'], cfg['modes']
    H = N + 256
    psi, probe, scan = synthetic.make_problem(sample, N, M, H, H, seed=0)
    data = onp.simulate(N, probe, scan, psi)
    psi0 = np.full_like(psi, 0.5 + 0j)
    return data, scan, psi0, probe


def time_cpu_oracle(cfg, sample, steps=1, warmup=0):
    from oracle import ptycho_np as onp
    data, scan, psi, probe = cpu_sample(cfg, sample)
    mask = np.ones(data.shape[-2:], bool)
    for _ in range(warmup):
        onp.rpie_batch(data, scan, psi, probe, mask)
    t0 = time.perf_counter()
    for _ in range(steps):
        onp.rpie_batch(data, scan, psi, probe, mask)
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cfg = dict(WORKLOAD)
    sample = 192
    value, dt = time_cpu_oracle(cfg, sample, steps=args.steps, warmup=min(args.warmup, 1))
    cores = os.cpu_count() or 1
    desc = (f'{sample} positions of the same workload (128x128, 8 modes) per step, '
            'forward + gradient math of rpie._get_nearplane_gradients')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'complex64', 'data': 'synthetic',
        'config': {'workload': 'rPIE 128x128 detector, 8 modes (BASELINE configs[1]), CPU sample',
                   **cfg},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'note': 'reference CuPy path cannot run (no CuPy in image); this is the '
                'NumPy/scipy.fft port of the reference algorithm on host cores',
    }))
    return 0


# ---------------------------------------------------------------- our arm ---
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tike_b200.ptycho as tp
    from tike_b200 import kernels as K

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    cfg = dict(WORKLOAD)
    if args.positions:
        cfg['positions_per_gpu'] = args.positions
    P_total = cfg['positions_per_gpu'] * world
    N = cfg['detector']

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
            e1.record()
            barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = K.launch_count() - launches0
        costs = [float(np.mean(c)) for c in ctx.parameters.algorithm_options.costs]

        # ------------ dominant kernel alone, for the roofline -------------
        p = ctx.parameters
        lo, hi = int(ctx.batches[0][0]), int(ctx.batches[0][-1]) + 1
        B = hi - lo
        batch = K.make_batch(p.psi[0], p.scan[lo:hi], p.probe[0, 0], N)
        kcost = torch.empty(B, device=device)
        psi_num = torch.zeros_like(p.psi)
        probe_num = torch.empty_like(p.probe[0, 0])
        kms = []
        for it in range(4):
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            K.rpie_batch(batch, ctx.data[lo:hi], None, N * N, noise_model='gaussian',
                         psi_numerator=psi_num[0], probe_numerator=probe_num,
                         costs=kcost, device=device)
            k1.record()
            torch.cuda.synchronize()
            if it:
                kms.append(k0.elapsed_time(k1))
        kernel_ms = statistics.mean(kms)
        del psi_num, probe_num

        # ------------ the reference's GPU op sequence with library ops -------
        standin = None
        if rank == 0 and not args.no_standin:
            from baseline import torch_standin
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
                       'sample': f'{nsamp} positions of batch 0 in 64-pattern chunks',
                       'what': 'reference rpie._get_nearplane_gradients op sequence restated '
                               'with torch.cuda library ops (cuFFT, gather, index_add); the '
                               'CuPy reference itself cannot run here (baseline/torch_standin.py)',
                       'fused_kernel_same_sample_ratio': None}
            standin['fused_kernel_same_sample_ratio'] = (B / (kernel_ms * 1e-3)) / standin['value']

    ms_per_step = ms_total / args.steps
    value = P_total / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API, host buffers ------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(data.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(data)
        torch.cuda.synchronize()
        del data
        torch.cuda.empty_cache()
        with tp.Reconstruction(host, params, split=split, data_is_local=True,
                               resident_data=False) as ctx:
            ctx.iterate(1)
            barrier()
            ksteps = max(1, min(args.steps, 3))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            ctx.iterate(ksteps)
            t1.record()
            barrier()
            e2e_ms = max_over_ranks(t0.elapsed_time(t1)) / ksteps
        e2e = {'value': P_total / (e2e_ms * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(host.numel() * 4 * world),
               'd2h_bytes_per_step': 4 * world, 'ms_per_step': e2e_ms,
               'steps': ksteps,
               'what': 'tike_b200.ptycho.Reconstruction.iterate(1) with pinned host '
                       'diffraction data re-streamed every epoch (resident_data=False)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'rpie_fast_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full
            # capture, scaled from its launch size to this launch's patterns
            traffic = json.load(f)['dram_bytes_per_pattern'] * B
    algo_bytes = ALGO_BYTES_PER_PATTERN * B
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu:
        v, dt = time_cpu_oracle(cfg, 192, steps=1, warmup=0)
        cpu = {'value': v, 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
               'sample': f'192 positions of the same workload, one rpie batch gradient '
                         f'pass of the NumPy/scipy.fft oracle ({dt:.1f} s)'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex64', 'data': 'synthetic',
        'config': {'workload': 'rPIE, 128x128 detector, 8 probe modes, 100k positions '
                               'per GPU, 4096x4096 complex64 object (BASELINE configs[1])',
                   **cfg, 'positions_total': P_total,
                   'l2': 'inputs exceed L2 (6.5 GB of patterns per GPU per epoch)',
                   'multi_gpu': 'replicated object/probe, NCCL all-reduce of numerators'},
        'clocks': clocks.summary(),
        'e2e': e2e,
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                     'traffic': traffic, 'peak_kind': peak_kind,
                     'kernel': 'rpie_fast_kernel<128>', 'kernel_ms': kernel_ms,
                     # the fused kernel sits above the FP32 ridge, so
                     # both components are reported; 22.8 MFLOP per pattern, FP32 peak
                     # 148 SM x 128 lanes x 2 x 1.965 GHz
                     'fp32': {'achieved_tflops': ALGO_FLOPS_PER_PATTERN * B / (kernel_ms * 1e-3) / 1e12,
                              'peak_tflops': FP32_PEAK_TFLOPS,
                              'frac': ALGO_FLOPS_PER_PATTERN * B / (kernel_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS},
                     'dram_frac': (traffic / (kernel_ms * 1e-3) / 1e9 / peaks['hbm_gbs']) if traffic else None,
                     'algorithmic_bytes': algo_bytes,
                     'patterns_per_launch': B,
                     'algorithmic_bytes_per_pattern': ALGO_BYTES_PER_PATTERN,
                     'note': 'fused kernel is FP32-issue / shared-memory / L2 bound, not HBM '
                             'bound (AI ~49 FLOP/B >> ridge 11); DRAM traffic above the '
                             'algorithmic bytes is the per-position far-field spill, '
                             'see DESIGN.md'},
        'cpu_baseline': cpu,
        'gpu_standin': standin,
        'cost_first_last': [costs[0], costs[-1]],
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
    ap.add_argument('--positions', type=int, default=0,
                    help='positions per GPU (default: 100000, the BASELINE config)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-standin', action='store_true',
                    help='skip the torch.cuda stand-in of the reference GPU path')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
