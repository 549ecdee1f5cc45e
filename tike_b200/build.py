"""Build libtikeb200.so (CUDA, sm_100a) in-tree with nvcc.

    python -m tike_b200.build [--force] [--verbose]

The shared library lands in tike_b200/lib/ so it travels with the repo
snapshot to the GPU box (a JIT cache under ~/.cache would not).
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(HERE, 'build')
LIBNAME = 'libtikeb200.so'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo', '--use_fast_math',
    '--expt-relaxed-constexpr',
    '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
]
# --use_fast_math would change sqrt/div accuracy: parity needs IEEE ops, so
# it is NOT used; the list above is rewritten below.
NVCC_FLAGS.remove('--use_fast_math')
# development switches, e.g. TB_NVCC_EXTRA=-DTB_PHASE_TIMING (see scripts/phase_timing.py)
NVCC_FLAGS += os.environ.get('TB_NVCC_EXTRA', '').split()


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'),
                 '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
               if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include',
                                'tike_b200.h'))
    objs, jobs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc, *NVCC_FLAGS, '-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f'--- nvcc {os.path.basename(s)}\n{r.stdout}{r.stderr}\n')
            else:
                # keep the register / spill report for DESIGN.md bookkeeping
                with open(os.path.join(OBJDIR, os.path.basename(s) + '.ptxas.log'), 'w') as f:
                    f.write(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f'nvcc failed on {s}')
    lib = os.path.join(LIBDIR, LIBNAME)
    if force or jobs or _stale(lib, objs):
        cmd = [nvcc, '-shared', '-o', lib, *objs, '-lcudart',
               '-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
