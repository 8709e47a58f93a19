"""Host-side multi-GPU logic: one process per GPU (torchrun), NCCL inside libxtpb200 for the data path.

The reference has no distributed layer on the GW-BSE path (one process, an OpenMP thread per GPU, host-side
reductions -- upstream xtp/src/libxtp/openmp_cuda.cc).  Here:

  * the RI tensor M[m][P][n] is distributed over its SECOND index, cyclically: rank r owns n = r, r + world, ...
    (occupied and unoccupied levels stay balanced); every stage that sums over n (epsilon, Sigma_x, Sigma_c diagonal
    and off-diagonal) produces a partial result that is all-reduced;
  * Fill3cMO splits its first half-transform over the auxiliary index (canonical contiguous ranges) and all-gathers
    the half-transformed blocks;
  * the BSE operator is distributed over the auxiliary index (contiguous ranges); one all-reduce of Y per matmul.

This module holds the partition arithmetic (shared with the CPU tests, which replay it with the numpy oracle over
gloo) and the communicator bootstrap through torch.distributed."""
from __future__ import annotations

import numpy as np


def aux_range(n_aux: int, rank: int, world: int):
    """Contiguous aux-function range of `rank` (Fill3cMO input slices, BSE operator shard)."""
    return n_aux * rank // world, n_aux * (rank + 1) // world


def local_columns(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global second-index positions (relative to nmin) held by `rank`: rank, rank + world, ..."""
    return np.arange(rank, n_total, world)


def n_local_below(g: int, rank: int, world: int) -> int:
    """Number of local columns with global position < g (e.g. g = n_occ gives the local occupied count)."""
    return (g - rank + world - 1) // world if g > rank else 0


def init_process_group_from_env(backend=None):
    """torch.distributed init from the torchrun environment; returns (rank, world, local_rank)."""
    import os

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def join_library_communicator(ctx, rank: int, world: int):
    """Rank 0 creates the library's NCCL id, torch.distributed broadcasts it, every rank joins."""
    if world == 1:
        return
    import torch.distributed as dist

    from . import api
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)
