"""cupyx.scipy -> scipy (test infrastructure only)."""
import contextlib
import sys
import types

import scipy.fft as _sfft
import scipy.ndimage as _snd
import scipy.stats as _sst

import cupy as _cp

fft = types.ModuleType('cupyx.scipy.fft')
for _n in dir(_sfft):
    if not _n.startswith('_'):
        _o = getattr(_sfft, _n)
        setattr(fft, _n, _cp._wrap(_o) if callable(_o) and not isinstance(_o, type) else _o)
fft.get_fft_plan = lambda a, axes=None, **k: contextlib.nullcontext()

ndimage = types.ModuleType('cupyx.scipy.ndimage')
for _n in dir(_snd):
    if not _n.startswith('_'):
        _o = getattr(_snd, _n)
        setattr(ndimage, _n, _cp._wrap(_o) if callable(_o) and not isinstance(_o, type) else _o)

stats = types.ModuleType('cupyx.scipy.stats')
for _n in dir(_sst):
    if not _n.startswith('_'):
        _o = getattr(_sst, _n)
        setattr(stats, _n, _cp._wrap(_o) if callable(_o) and not isinstance(_o, type) else _o)

for _m in (fft, ndimage, stats):
    sys.modules[_m.__name__] = _m
