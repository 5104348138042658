"""Data-parallel plumbing: one process per GPU, torch.distributed for rendezvous and the
gradient all-reduce (NCCL over NVLink 5 / NVSwitch on GPUs, gloo on CPU for host-logic tests).

The path shards naturally over structures (no cross-structure op exists in the graph,
SURVEY.md 8e).  The only exchange is ONE sum all-reduce per train step of the flat gradient
arena with the batch SSE appended: the 1/(B*RMSE) factor of the non-additive RMSE loss
(scann/layers/losses.py:5-6) is applied after the reduce, inside the optimiser kernel.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def allreduce_sum(t: torch.Tensor) -> None:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n structures for this rank (SURVEY.md 8e partitioning)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def attach(model, world: int) -> None:
    """Make ``model.train_on_batch`` data-parallel: local shard in, global-batch semantics out."""
    if world > 1:
        model.allreduce = allreduce_sum
        model.world_size = world
        # every rank draws its own Dropout masks (its shard holds different structures)
        rank = env_world()[0]
        model.engine.dropout_seed = (model.engine.dropout_seed + 7919 * rank) & 0x7FFFFFFF
