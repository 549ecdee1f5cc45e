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
