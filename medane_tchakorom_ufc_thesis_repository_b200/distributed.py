"""One process per GPU: plumbing between ranks with torch.distributed.

The data path never goes through torch: boundary layers move by P2P stores into CUDA-IPC mapped
peer windows and the scalar / TSQR / flag reductions run on the engine's own NCCL communicator
(include/msplit.h, section "multi-block (b)").  torch.distributed only carries the bootstrap
bytes (NCCL unique id, IPC handles) and the max-over-ranks of the measured times.  Everything in
this file also works with the ``gloo`` backend on CPU, which is how the CPU test-suite covers it.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def block_of_rank(rank: int, nprocs: int, npb: int = 1) -> Tuple[int, int]:
    """computeDimensionRelatedVariables utils.c:657-659: (rank_jacobi_block, proc_local_rank).
    The GPU engine runs one process per block (npb = 1)."""
    if npb < 1 or nprocs % npb:
        raise ValueError("nprocs must be a multiple of npb")
    return rank // npb, rank % npb


def strip_partition(layers: int, nblocks: int) -> List[Tuple[int, int]]:
    """1-D strip partition of the grid lines (2-D) / z planes (3-D): block K owns [K*L/G, (K+1)*L/G)
    (the reference's rule at G = 2, utils.c:254-264)."""
    if nblocks < 1 or layers % nblocks:
        raise ValueError("grid lines / planes must be divisible by the number of blocks")
    per = layers // nblocks
    return [(k * per, (k + 1) * per) for k in range(nblocks)]


def neighbours(block: int, nblocks: int) -> List[Optional[int]]:
    """[lower, upper] neighbour block of a strip (None at the domain boundary)."""
    return [block - 1 if block > 0 else None, block + 1 if block < nblocks - 1 else None]


def init_process_group(backend: Optional[str] = None):
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world == 1:
        return None
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return dist


def allgather_bytes(payload: bytes) -> List[bytes]:
    """Every rank contributes a byte string; returns all of them in rank order."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return [payload]
    out: List[Optional[bytes]] = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return [bytes(o) for o in out]


def broadcast_bytes(payload: Optional[bytes], src: int = 0) -> bytes:
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        assert payload is not None
        return payload
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return bytes(box[0])


def reduce_max(value: float) -> float:
    """max over ranks (multi-GPU times are reported as the max over ranks of device-timed regions)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def connect_blocks(rank: int, world: int, export_window: Callable[[], bytes], connect: Callable[[int, bytes], None],
                   make_unique_id: Callable[[], bytes], comm_init: Callable[[bytes, int, int], None],
                   connect_block: Optional[Callable[[int, bytes], None]] = None) -> None:
    """Bootstrap of the per-GPU engines: rank 0 creates the NCCL id, every rank joins the communicator,
    exports its receive window and maps the windows of blocks K-1 / K+1 (replaces the MPI communicators
    built at …multisplitting.c:66-77)."""
    uid = broadcast_bytes(make_unique_id() if rank == 0 else None, 0)
    comm_init(uid, rank, world)
    handles = allgather_bytes(export_window())
    if connect_block is not None:
        for blk in range(world):
            if blk != rank:
                connect_block(blk, handles[blk])
    else:
        for side, nb in enumerate(neighbours(rank, world)):
            if nb is not None:
                connect(side, handles[nb])
    barrier()


def make_distributed_engine(m, n, p=1, s=0, max_restart=30, keep_csr=False, npb=1):
    """Engine of this rank's strip, wired to its neighbours.  Call under torchrun (one rank per GPU).
    npb = GPUs per Jacobi block (the reference's -npb): Jacobi blocks = WORLD_SIZE / npb, rank r belongs to block r // npb."""
    from . import solver
    rank, world, local = env_rank()
    init_process_group()
    if world % npb:
        raise ValueError("WORLD_SIZE must be a multiple of npb")
    eng = solver.Engine(m, n, p, block=rank, nblocks=world, s=s, max_restart=max_restart, device=local, keep_csr=keep_csr, npb=npb)
    if world > 1:
        connect_blocks(rank, world, eng.comm_export, eng.comm_connect, solver.comm_unique_id, eng.comm_init,
                       eng.comm_connect_block)
    return eng
