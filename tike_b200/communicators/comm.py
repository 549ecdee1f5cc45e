"""torch.distributed-backed communicator."""
from __future__ import annotations

import typing

import numpy as np
import torch
import torch.distributed as dist


class Comm:
    """Collectives used by the reconstruction driver.

    ``size == 1`` (no process group) makes every method a no-op / identity so
    single-GPU runs need no rendezvous.  Reference call sites:
    ptycho.py:474-515 (probe mean, psi halo swap, transform mean, cost
    gather), _preconditioner.py:185,201 (commented-out Allreduce) and
    ptycho.py:946 (init rescale reduce).
    """

    def __init__(self, group=None):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.size = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.size = 1
            self.rank = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # -- in-place tensor collectives ------------------------------------
    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            buf = torch.view_as_real(t) if t.is_complex() else t
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_mean_(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            self.allreduce_sum_(t)
            t /= self.size
        return t

    def bcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.size > 1:
            buf = torch.view_as_real(t) if t.is_complex() else t
            dist.broadcast(buf, src=src, group=self.group)
        return t

    def barrier(self):
        if self.size > 1:
            dist.barrier(group=self.group)

    # -- small host objects ----------------------------------------------
    def allgather_object(self, obj) -> list:
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def bcast_object(self, obj, src: int = 0):
        if self.size == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def reduce_cpu_sum(self, x: np.ndarray) -> np.ndarray:
        """Sum of small host arrays over ranks (ptycho.py:946)."""
        if self.size == 1:
            return x
        return np.sum(self.allgather_object(np.asarray(x)), axis=0)

    # -- stripe mode (reference-faithful halo blend) ---------------------
    def swap_edges(self, psi: torch.Tensor, overlap: int,
                   edges: typing.Sequence[int]) -> torch.Tensor:
        """Blend the band [edge, edge + overlap) of neighbouring stripes with
        linear ramps (pool.py:415-476) — rank i exchanges with i+1."""
        if self.size == 1:
            return psi
        if overlap < 1:
            raise ValueError(
                f"Overlap for swap_edges cannot be less than 1: {overlap}")
        for i in range(self.size - 1):
            lo = edges[i + 1]
            hi = lo + overlap
            if self.rank not in (i, i + 1):
                continue
            mine = psi[..., lo:hi, :].contiguous()
            theirs = torch.empty_like(mine)
            peer = i + 1 if self.rank == i else i
            a, b = torch.view_as_real(mine), torch.view_as_real(theirs)
            ops = [dist.P2POp(dist.isend, a, peer, self.group),
                   dist.P2POp(dist.irecv, b, peer, self.group)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            lower, upper = (mine, theirs) if self.rank == i else (theirs, mine)
            psi[..., lo:hi, :] = swap_edges_pair(lower, upper, overlap,
                                                 is_lower=self.rank == i)
        return psi


def swap_edges_pair(lower, upper, overlap: int, is_lower: bool):
    """Blended band for the lower (i) or upper (i+1) stripe owner:
    lower gets rampd*lower + rampu*upper, upper gets the same blend
    (pool.py:446-475)."""
    ramp = torch.linspace(0.0, 1.0, overlap + 2, device=lower.device,
                          dtype=torch.float32)[1:-1][..., None]
    return (1.0 - ramp) * lower + ramp * upper


def stitch_stripes(parts, stripe_start, probe_width: int):
    """Host-side stitch of per-rank objects (ObjectOptions.join_psi,
    object.py:154-167): rank i owns the rows from the centre of its first
    probe footprint to the centre of the next rank's; rank 0 also keeps
    everything above, the last rank everything below."""
    out = parts[0]
    half = probe_width // 2
    rows = out.shape[1]
    cuts = [int(s) + half for s in stripe_start[1:len(parts)]] + [rows]
    for i in range(1, len(parts)):
        out[:, cuts[i - 1]:cuts[i], :] = parts[i][:, cuts[i - 1]:cuts[i], :]
    return out
