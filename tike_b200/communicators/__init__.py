"""Multi-GPU communication: one process per GPU over torch.distributed
(NCCL on GPUs, gloo in CPU tests).

Replaces the reference's thread-per-GPU ``Comm`` / ``ThreadPool`` / ``MPIComm``
stack (src/tike/communicators/{comm,pool,mpi}.py), whose "collectives" are
serial peer copies issued from Python threads (pool.py:300-395).
"""
from .comm import Comm, RowPlan, swap_edges_pair, stitch_stripes

__all__ = ['Comm', 'RowPlan', 'swap_edges_pair', 'stitch_stripes']
