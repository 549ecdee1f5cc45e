"""ctypes binding of libtikeb200.so (the C-ABI declared in include/tike_b200.h).

Arrays cross the boundary as raw device pointers read from
``__cuda_array_interface__`` (torch CUDA tensors, CuPy arrays, Numba device
arrays ... anything that exports the protocol).  There is no CPU fallback: if
the shared library is missing, importing the compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TB_LIB_PATH: development switch, load an experimental build of the library
# (scripts/build_variant.py) instead of the in-tree one
LIB_PATH = os.environ.get('TB_LIB_PATH') or os.path.join(_HERE, 'lib', 'libtikeb200.so')

_lib = None
_lock = threading.Lock()


class LibraryNotBuilt(RuntimeError):
    pass


class tb_batch(C.Structure):
    _fields_ = [
        ('psi', C.c_void_p), ('height', C.c_int32), ('width', C.c_int32),
        ('scan', C.c_void_p), ('npos', C.c_int64),
        ('probe', C.c_void_p), ('nmodes', C.c_int32), ('probe_width', C.c_int32),
        ('probe_per_position', C.c_int32),
        ('eigen_probe', C.c_void_p), ('neigen', C.c_int32), ('eigen_modes', C.c_int32),
        ('eigen_weights', C.c_void_p),
        ('detector_width', C.c_int32),
        ('fwd_scale', C.c_float), ('inv_scale', C.c_float),
    ]


class tb_rpie_args(C.Structure):
    _fields_ = [
        ('batch', tb_batch),
        ('data', C.c_void_p), ('data_dtype', C.c_int32),
        ('mask', C.c_void_p), ('num_measured', C.c_int32),
        ('noise_model', C.c_int32), ('step_mode', C.c_int32),
        ('step_length_start', C.c_float), ('step_length_weight', C.c_float),
        ('unmeasured_scaling', C.c_float),
        ('accumulate_object', C.c_int32),
        ('psi_numerator', C.c_void_p), ('probe_numerator', C.c_void_p),
        ('costs', C.c_void_p), ('eigen_weight_step', C.c_void_p),
        ('workspace', C.c_void_p), ('workspace_bytes', C.c_int64),
    ]


class tb_lstsq_args(C.Structure):
    _fields_ = [
        ('batch', tb_batch),
        ('data', C.c_void_p), ('data_dtype', C.c_int32),
        ('mask', C.c_void_p), ('num_measured', C.c_int32),
        ('noise_model', C.c_int32), ('step_mode', C.c_int32),
        ('step_length_start', C.c_float), ('step_length_weight', C.c_float),
        ('unmeasured_scaling', C.c_float),
        ('recover_psi', C.c_int32), ('recover_probe', C.c_int32),
        ('recover_positions', C.c_int32),
        ('chi', C.c_void_p), ('object_upd_sum', C.c_void_p),
        ('probe_upd_sum', C.c_void_p), ('costs', C.c_void_p),
        ('position_num', C.c_void_p), ('position_den', C.c_void_p),
        ('gradient_taps', C.c_float * 5),
        ('workspace', C.c_void_p), ('workspace_bytes', C.c_int64),
    ]


EXPORTS = [
    'tb_last_error', 'tb_version', 'tb_sm_count', 'tb_patch_fwd',
    'tb_patch_adj', 'tb_fft2', 'tb_ptycho_fwd', 'tb_rpie_workspace_size',
    'tb_rpie_batch', 'tb_rpie_update_psi', 'tb_rpie_update_probe',
    'tb_precond_psi', 'tb_precond_probe', 'tb_lstsq_workspace_size',
    'tb_lstsq_phase1', 'tb_lstsq_phase2', 'tb_lstsq_precondition_object',
    'tb_caxpy', 'tb_cluster_grow', 'tb_cluster_compact_sweep', 'tb_multislice_workspace_size',
    'tb_multislice_fwd', 'tb_multislice_rpie_batch', 'tb_multislice_precond_psi',
    'tb_affine_inliers', 'tb_max_real', 'tb_rpie_update_psi_given_max',
    'tb_rpie_update_psi_adam', 'tb_momentum_update',
    'tb_lstsq_precondition_object_given_max', 'tb_add_quotient',
    'tb_object_pointwise_constraints', 'tb_object_smoothness',
    'tb_weighted_norm_sums', 'tb_scale_by_device_scalar', 'tb_multislice_lstsq_phase1',
    'tb_lstsq_eigen_pass1', 'tb_lstsq_eigen_pass2',
]


def lib():
    """Load (once) and return the ctypes handle of libtikeb200.so."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise LibraryNotBuilt(
                f'{LIB_PATH} not found: build it with `python -m tike_b200.build` '
                '(needs nvcc; there is no CPU fallback).')
        h = C.CDLL(LIB_PATH)
        vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
        h.tb_last_error.restype = C.c_char_p
        h.tb_last_error.argtypes = []
        h.tb_version.restype = i32
        h.tb_sm_count.argtypes = [C.POINTER(C.c_int)]
        h.tb_patch_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
        h.tb_patch_adj.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
        h.tb_fft2.argtypes = [vp, i64, i32, i32, f32, vp]
        h.tb_ptycho_fwd.argtypes = [C.POINTER(tb_batch), vp, vp, vp]
        h.tb_rpie_workspace_size.argtypes = [C.POINTER(tb_rpie_args)]
        h.tb_rpie_workspace_size.restype = i64
        h.tb_rpie_batch.argtypes = [C.POINTER(tb_rpie_args), vp]
        h.tb_rpie_update_psi.argtypes = [vp, vp, vp, i64, f32, vp, vp]
        h.tb_rpie_update_probe.argtypes = [vp, vp, vp, i32, i64, f32, vp, vp]
        h.tb_precond_psi.argtypes = [vp, i32, i32, vp, vp, i64, vp, i32, i32, vp, vp]
        h.tb_precond_probe.argtypes = [vp, i32, i32, vp, vp, i64, i32, vp, vp]
        h.tb_lstsq_workspace_size.argtypes = [C.POINTER(tb_lstsq_args)]
        h.tb_lstsq_workspace_size.restype = i64
        h.tb_lstsq_phase1.argtypes = [C.POINTER(tb_lstsq_args), vp]
        h.tb_lstsq_phase2.argtypes = [C.POINTER(tb_batch), vp, vp, vp, i32, f32, vp, vp]
        h.tb_lstsq_precondition_object.argtypes = [vp, vp, vp, i64, f32, vp, vp]
        h.tb_caxpy.argtypes = [vp, vp, i64, f32, vp, vp]
        h.tb_cluster_grow.argtypes = [vp, i64, i32, vp, i32, i64]
        h.tb_cluster_compact_sweep.argtypes = [vp, vp, vp, vp, vp, i64, i32]
        h.tb_affine_inliers.argtypes = [vp, vp, vp, vp, i64, vp, C.c_double, C.c_double,
                                        C.c_double, vp, C.POINTER(C.c_int64)]
        h.tb_max_real.argtypes = [vp, i64, vp, vp]
        h.tb_rpie_update_psi_given_max.argtypes = [vp, vp, vp, i64, f32, vp, vp]
        h.tb_rpie_update_psi_adam.argtypes = [vp, vp, vp, vp, vp, i64, f32, C.c_double, C.c_double, vp, vp]
        h.tb_momentum_update.argtypes = [vp, vp, vp, i64, C.c_double, vp, vp]
        h.tb_lstsq_precondition_object_given_max.argtypes = [vp, vp, vp, i64, f32, vp, vp]
        h.tb_add_quotient.argtypes = [vp, vp, vp, i64, i64, f32, vp]
        h.tb_object_pointwise_constraints.argtypes = [vp, i64, f32, i32, f32, vp]
        h.tb_object_smoothness.argtypes = [vp, vp, i32, i32, i32, f32, vp]
        h.tb_weighted_norm_sums.argtypes = [vp, vp, i64, vp, vp]
        h.tb_scale_by_device_scalar.argtypes = [vp, i64, vp, i32, vp]
        h.tb_multislice_workspace_size.argtypes = [C.POINTER(tb_batch), i32]
        h.tb_multislice_workspace_size.restype = i64
        h.tb_multislice_fwd.argtypes = [C.POINTER(tb_batch), i32, vp, vp, vp, vp, i64, vp]
        h.tb_multislice_rpie_batch.argtypes = [C.POINTER(tb_rpie_args), i32, vp, vp]
        h.tb_multislice_precond_psi.argtypes = [C.POINTER(tb_batch), i32, vp, vp, vp, i64, vp]
        h.tb_multislice_lstsq_phase1.argtypes = [C.POINTER(tb_lstsq_args), i32, vp, vp]
        h.tb_lstsq_eigen_pass1.argtypes = [C.POINTER(tb_batch), vp, i32, vp, vp, i64, i32, vp, i32,
                                           vp, i64, vp, vp, vp, vp]
        h.tb_lstsq_eigen_pass2.argtypes = [C.POINTER(tb_batch), vp, i32, vp, vp, i64, i32, vp, i32,
                                           vp, vp, vp]
        for name in EXPORTS:
            f = getattr(h, name)
            if name not in ('tb_last_error', 'tb_rpie_workspace_size',
                            'tb_lstsq_workspace_size', 'tb_multislice_workspace_size'):
                f.restype = i32
        _lib = h
    return _lib


def check(rc: int, what: str = ''):
    """Raise like the reference would: ValueError for bad arguments/shapes,
    RuntimeError for CUDA failures."""
    if rc == 0:
        return
    msg = lib().tb_last_error().decode(errors='replace')
    if rc in (-1, -2):
        raise ValueError(f'{what}: {msg}' if what else msg)
    raise RuntimeError(f'{what}: CUDA error {rc}: {msg}')


_TYPESTR = {
    '<c8': np.complex64, '<f4': np.float32, '<u2': np.uint16, '|u1': np.uint8,
    '|b1': np.bool_, '<i4': np.int32, '<f8': np.float64,
}


def dev_ptr(x, typestr=None, name='array') -> int:
    """Device pointer of a C-contiguous ``__cuda_array_interface__`` exporter."""
    if x is None:
        return None
    try:
        cai = x.__cuda_array_interface__
    except AttributeError:
        raise TypeError(f'{name} must export __cuda_array_interface__ '
                        f'(got {type(x).__name__}); host arrays are not accepted '
                        'by the compute path') from None
    if typestr is not None:
        allowed = (typestr,) if isinstance(typestr, str) else tuple(typestr)
        if cai['typestr'] not in allowed:
            raise ValueError(f'{name} has dtype {cai["typestr"]}, expected {allowed}')
    strides = cai.get('strides')
    if strides is not None:
        shape = cai['shape']
        item = int(cai['typestr'][2:])
        expect = []
        acc = item
        for n in reversed(shape):
            expect.append(acc)
            acc *= max(int(n), 1)
        expect = tuple(reversed(expect))
        if any(int(n) > 1 and int(s) != e for n, s, e in zip(shape, strides, expect)):
            raise ValueError(f'{name} must be C-contiguous')
    return int(cai['data'][0])


def stream_ptr(stream=None) -> int:
    """cudaStream_t of ``stream`` (torch / cupy stream object or raw int);
    default: torch's current stream, which is what a caller inside
    ``with torch.cuda.stream(s):`` expects (reference: pool.py:406-409)."""
    if stream is None:
        import torch
        return int(torch.cuda.current_stream().cuda_stream)
    if isinstance(stream, int):
        return stream
    for attr in ('cuda_stream', 'ptr'):
        if hasattr(stream, attr):
            return int(getattr(stream, attr))
    raise TypeError(f'cannot get a cudaStream_t from {type(stream).__name__}')


def sm_count() -> int:
    n = C.c_int(0)
    check(lib().tb_sm_count(C.byref(n)), 'tb_sm_count')
    return n.value
