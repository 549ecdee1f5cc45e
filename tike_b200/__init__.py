"""tike_b200 — a B200-native (sm_100a) implementation of the ptychography
reconstruction hot path of AdvancedPhotonSource/tike.

Python host code mirrors the reference API (``tike_b200.ptycho.reconstruct``,
``simulate``, ``PtychoParameters``, ``RpieOptions``, ``LstsqOptions``,
``DmOptions``, the operator classes); the arithmetic runs in hand-written CUDA
kernels behind a C ABI (include/tike_b200.h, tike_b200/csrc).  Submodules are
imported lazily so that ``import tike_b200`` stays cheap.
"""
import importlib

__version__ = '0.1.0'

_SUBMODULES = ('ptycho', 'operators', 'cluster', 'communicators', 'constants', 'kernels',
               'linalg', 'opt', 'precision', 'random', 'synthetic', 'build')


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module(f'{__name__}.{name}')
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
