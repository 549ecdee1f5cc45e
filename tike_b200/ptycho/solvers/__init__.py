"""Iterative ptychography solvers (reference:
src/tike/ptycho/solvers/__init__.py:1-16, plus ``dm`` / ``DmOptions``)."""
from .options import (IterativeOptions, RpieOptions, LstsqOptions, DmOptions,
                      PtychoParameters, crop_fourier_space, ExitWaveOptions,
                      ObjectOptions, PositionOptions, ProbeOptions)
from ._preconditioner import update_preconditioners
from .lstsq import lstsq_grad
from .rpie import rpie
from .dm import dm

__all__ = [
    'crop_fourier_space', 'lstsq_grad', 'rpie', 'dm', 'LstsqOptions',
    'RpieOptions', 'DmOptions', 'PtychoParameters', 'update_preconditioners',
]
