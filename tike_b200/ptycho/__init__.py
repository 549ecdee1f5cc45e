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


def _reexport_public_names():
    """The reference's package does ``from .module import *`` for every
    submodule (ptycho/__init__.py:2-8), so ``tike.ptycho.<function>`` works for
    every public helper; mirror that for the functions and classes defined in
    our submodules."""
    import inspect
    from . import ptycho as _ptycho
    for mod in (object, position, probe, exitwave, _ptycho, solvers):
        for name, value in vars(mod).items():
            if name.startswith('_') or name in globals():
                continue
            if (inspect.isfunction(value) or inspect.isclass(value)) and \
                    getattr(value, '__module__', '').startswith(__name__):
                globals()[name] = value


_reexport_public_names()
