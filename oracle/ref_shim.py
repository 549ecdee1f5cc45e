"""Load the UNMODIFIED reference package on a CPU-only host (TEST INFRA ONLY).

``load_reference()`` puts oracle/refshim (a NumPy-backed ``cupy``/``cupyx``)
and /root/reference/src on ``sys.path``, imports ``tike`` and replaces the two
pieces that need a real GPU:

* ``tike.operators.cupy.patch.Patch.fwd/adj`` (NVRTC RawModule kernels,
  patch.py:35-39) -> oracle.ptycho_np.patch_fwd/patch_adj, the NumPy
  restatement of convolution.cu which is itself pinned by the reference's
  known-answer tests (tests/test_oracle.py);
* ``tike.communicators.stream.stream_and_modify2`` (CUDA streams) -> the
  reference's own ``stream_and_modify_debug2`` (stream.py:407-458).

It only works where /root/reference exists (the build container); it is used
by tests/golden/make_golden.py to produce the committed fixtures and by
``-m "not gpu"`` tests that skip when the reference is absent.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REFERENCE_SRC = os.environ.get('TIKE_REFERENCE_SRC', '/root/reference/src')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'refshim')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'tike'))


def load_reference():
    """Return the imported reference ``tike`` module running on the shim."""
    if not reference_available():
        raise RuntimeError(f'reference not found at {REFERENCE_SRC}')
    if 'tike' in sys.modules and getattr(sys.modules['tike'], '_on_shim', False):
        return sys.modules['tike']
    for p in (REFERENCE_SRC, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cupy  # noqa: F401  (the shim)
    import cupyx  # noqa: F401
    import tike
    import tike.ptycho
    import importlib
    refpatch = importlib.import_module('tike.operators.cupy.patch')
    refstream = importlib.import_module('tike.communicators.stream')
    from . import ptycho_np as onp

    def _fwd(self, images, positions, patches=None, patch_width=0, height=0,
             width=0, nrepeat=1):
        patch_width = patches.shape[-1] if patch_width == 0 else patch_width
        lead = positions.shape[:-2]
        assert images.shape[:-2] == lead
        if patches is None:
            patches = cupy.zeros(
                (*lead, positions.shape[-2] * nrepeat, patch_width,
                 patch_width), dtype=images.dtype)
        assert positions.shape[-2] * nrepeat == patches.shape[-3]
        im = np.asarray(images).reshape(-1, *images.shape[-2:])
        po = np.asarray(positions).reshape(-1, *positions.shape[-2:])
        pa = np.asarray(patches).reshape(-1, *patches.shape[-3:])
        for i in range(im.shape[0]):
            onp.patch_fwd(im[i], po[i], patch_width, nrepeat=nrepeat,
                          patches=pa[i])
        return patches

    def _adj(self, positions, patches, images=None, patch_width=0, height=0,
             width=0, nrepeat=1):
        patch_width = patches.shape[-1] if patch_width == 0 else patch_width
        lead = positions.shape[:-2]
        if images is None:
            images = cupy.zeros((*lead, height, width), dtype=patches.dtype)
        im = np.asarray(images).reshape(-1, *images.shape[-2:])
        po = np.asarray(positions).reshape(-1, *positions.shape[-2:])
        pa = np.asarray(patches).reshape(-1, *patches.shape[-3:])
        for i in range(im.shape[0]):
            onp.patch_adj(po[i], pa[i], im[i], patch_width, nrepeat=nrepeat)
        return images

    refpatch.Patch.fwd = _fwd
    refpatch.Patch.adj = _adj

    # Worker-to-worker copies: on real GPUs ThreadPool._copy_to (pool.py:114-120)
    # is cupy.asarray under another device, i.e. a peer COPY.  The shim has one
    # address space, where asarray would alias the source and swap_edges
    # (pool.py:441-475) would read an already blended halo.  Keep copy semantics.
    refpool = importlib.import_module('tike.communicators.pool')

    def _copy_to(self, x, worker):
        return cupy.array(x, copy=True)

    refpool.ThreadPool._copy_to = _copy_to
    refstream.stream_and_modify2 = refstream.stream_and_modify_debug2
    tike.communicators.stream.stream_and_modify2 = refstream.stream_and_modify_debug2
    tike._on_shim = True
    return tike


def seed_reference(tike, seed: int):
    """Seed the two global generators the reference draws batch order from
    (random.py:10; cluster.py:518-532 uses legacy np.random)."""
    import tike.random
    tike.random.randomizer_np = np.random.default_rng(seed)
    np.random.seed(seed)
