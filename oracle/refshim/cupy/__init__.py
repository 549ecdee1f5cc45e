"""NumPy-backed stand-in for the `cupy` package (TEST INFRASTRUCTURE ONLY).

This module exists so that the *unmodified* reference package under
/root/reference/src can be imported and executed on a CPU-only container in
order to generate golden vectors (tests/golden/make_golden.py) and to validate
the NumPy restatement in oracle/ptycho_np.py.  It is never imported by the
product package `tike_b200`.

Every public NumPy callable is re-exported through a thin wrapper that
  * converts ``axis=[...]`` lists to tuples (reference: rpie.py:376-379), and
  * returns results as an ``ndarray`` subclass that has ``.get()``/``.set()``
    and keeps 0-d results as arrays (CuPy never returns Python scalars).
"""
import builtins
import contextlib
import functools
import types

import numpy as _np

from . import cuda  # noqa: F401


class ndarray(_np.ndarray):
    """numpy array that quacks like cupy.ndarray."""

    def get(self, *a, **k):
        return _np.array(self.view(_np.ndarray), copy=True)

    def set(self, x):
        self[...] = x

    @property
    def device(self):
        return cuda.Device(0)

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        # keep 0-d results as arrays, like CuPy does
        return _np.asarray(arr).view(ndarray)

    def __array_finalize__(self, obj):
        pass

    def item(self, *a):
        return self.view(_np.ndarray).item(*a)


def _wrap_result(r):
    if isinstance(r, ndarray):
        return r
    if isinstance(r, _np.ndarray):
        return r.view(ndarray)
    if isinstance(r, _np.generic):
        return _np.asarray(r).view(ndarray)
    if isinstance(r, tuple):
        return tuple(_wrap_result(x) for x in r)
    if isinstance(r, list):
        return [_wrap_result(x) for x in r]
    return r


def _wrap(f):
    @functools.wraps(f)
    def g(*args, **kwargs):
        if isinstance(kwargs.get('axis', None), list):
            kwargs['axis'] = tuple(kwargs['axis'])
        return _wrap_result(f(*args, **kwargs))
    return g


def _export(namespace, module):
    for name in dir(module):
        if name.startswith('_') or name in ('ndarray',):
            continue
        obj = getattr(module, name)
        if isinstance(obj, (types.FunctionType, types.BuiltinFunctionType,
                            _np.ufunc)) or (callable(obj)
                                            and not isinstance(obj, type)):
            namespace[name] = _wrap(obj)
        else:
            namespace[name] = obj


_export(globals(), _np)


def _namespace(module):
    ns = types.SimpleNamespace()
    d = {}
    _export(d, module)
    for k, v in d.items():
        setattr(ns, k, v)
    return ns


fft = _namespace(_np.fft)
linalg = _namespace(_np.linalg)
random = _namespace(_np.random)

# dtypes / classes that must stay classes
for _n in ('float32', 'float64', 'complex64', 'complex128', 'single', 'double',
           'csingle', 'cdouble', 'intc', 'int32', 'int64', 'uint16', 'uint8',
           'bool_', 'dtype', 'newaxis', 'pi', 'inf', 'nan'):
    globals()[_n] = getattr(_np, _n)


def asarray(a, dtype=None, order=None, **kw):
    return _np.asarray(a, dtype=dtype, order=order).view(ndarray)


def array(a, dtype=None, copy=True, **kw):
    return _np.array(a, dtype=dtype, copy=copy).view(ndarray)


def asnumpy(a, *args, **kw):
    if a is None:
        return None
    return _np.asarray(a).view(_np.ndarray)


def get_array_module(*args):
    import sys
    return sys.modules[__name__]


def fuse(*a, **k):
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]

    def deco(f):
        return f
    return deco


class RawModule:
    def __init__(self, *a, **k):
        pass

    def get_function(self, name):
        raise RuntimeError('RawModule kernels are not available in the shim')


class _Pool:
    def free_all_blocks(self):
        pass

    def used_bytes(self):
        return 0

    def total_bytes(self):
        return 0


def get_default_memory_pool():
    return _Pool()


def get_default_pinned_memory_pool():
    return _Pool()


def zeros_like(a, dtype=None, shape=None, **kw):
    return _np.zeros_like(_np.asarray(a), dtype=dtype, shape=shape).view(ndarray)


def empty_like(a, dtype=None, shape=None, **kw):
    # zeros: deterministic, and the reference relies on fills anyway
    return _np.zeros_like(_np.asarray(a), dtype=dtype, shape=shape).view(ndarray)


def ones_like(a, dtype=None, shape=None, **kw):
    return _np.ones_like(_np.asarray(a), dtype=dtype, shape=shape).view(ndarray)


def full_like(a, fill_value, dtype=None, shape=None, **kw):
    return _np.full_like(_np.asarray(a), fill_value, dtype=dtype,
                         shape=shape).view(ndarray)


def empty(shape, dtype=float, **kw):
    return _np.zeros(shape, dtype=dtype).view(ndarray)


def unravel_index(indices, dims=None, order='C', shape=None):
    """cupy.unravel_index names the shape argument ``dims``."""
    return tuple(_np.asarray(x).view(ndarray) for x in
                 _np.unravel_index(indices, dims if dims is not None else shape, order=order))
