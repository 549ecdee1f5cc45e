"""Forward / adjoint operators of the ptychography path (reference:
src/tike/operators/cupy/).  Same names, keyword arguments and attributes as
the reference; the arithmetic runs in libtikeb200 (CUDA, sm_100a)."""
from .operator import Operator
from .patch import Patch
from .convolution import Convolution
from .propagation import FresnelSpectProp, Propagation, ZeroPropagation
from .ptycho import Ptycho, Multislice, SingleSlice
from .objective import (gaussian, gaussian_grad, gaussian_each_pattern,
                        poisson, poisson_grad, poisson_each_pattern)

__all__ = [
    'Operator', 'Patch', 'Convolution', 'Propagation', 'ZeroPropagation',
    'FresnelSpectProp',
    'Ptycho', 'Multislice', 'SingleSlice', 'gaussian', 'gaussian_grad',
    'gaussian_each_pattern', 'poisson', 'poisson_grad', 'poisson_each_pattern',
]
