"""Random number helpers (reference: src/tike/random.py:10-26).

``randomizer_np`` is the generator the solvers draw the per-epoch batch order
from (rpie.py:95-98, lstsq.py:88-91); tests replace it with a seeded generator
to reproduce a reference trajectory.
"""
import numpy as np

from . import precision

randomizer_np = np.random.default_rng()


def numpy_complex(*shape):
    """Complex random array with real and imaginary parts in [-0.5, 0.5)."""
    pair = randomizer_np.random(size=(*shape, 2), dtype=precision.floating) - 0.5
    return pair.view(precision.cfloating)[..., 0]


def cupy_complex(*shape):
    """Device counterpart of numpy_complex (random.py:22-26): a complex64 CUDA
    tensor with real and imaginary parts in [-0.5, 0.5)."""
    import torch
    pair = torch.rand((*shape, 2), dtype=torch.float32, device='cuda') - 0.5
    return torch.view_as_complex(pair)


def cluster_wobbly_center(*args, **kwargs):
    """Deprecated alias of cluster.wobbly_center (random.py:29-38)."""
    import warnings
    warnings.warn('random.cluster_wobbly_center is deprecated. '
                  'Use cluster.wobbly_center instead.', DeprecationWarning)
    from . import cluster
    return cluster.wobbly_center(*args, **kwargs)


def cluster_compact(*args, **kwargs):
    """Deprecated alias of cluster.compact (random.py:41-50)."""
    import warnings
    warnings.warn('random.cluster_compact is deprecated. '
                  'Use cluster.compact instead.', DeprecationWarning)
    from . import cluster
    return cluster.compact(*args, **kwargs)
