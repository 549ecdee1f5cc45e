"""Ptychography reconstruction (reference: src/tike/ptycho/__init__.py)."""
from . import exitwave, object, position, probe, solvers
from .exitwave import ExitWaveOptions
from .object import ObjectOptions, get_padded_object
from .position import PositionOptions, AffineTransform, check_allowed_positions
from .probe import ProbeOptions
from .solvers import (PtychoParameters, RpieOptions, LstsqOptions, DmOptions,
                      IterativeOptions)
from .ptycho import (reconstruct, simulate, Reconstruction,
                     reconstruct_multigrid)

__all__ = [
    'reconstruct', 'simulate', 'Reconstruction', 'reconstruct_multigrid',
    'PtychoParameters', 'RpieOptions', 'LstsqOptions', 'DmOptions',
    'ExitWaveOptions', 'ObjectOptions', 'PositionOptions', 'ProbeOptions',
    'AffineTransform', 'check_allowed_positions', 'get_padded_object',
]
