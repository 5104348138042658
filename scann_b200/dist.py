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


_native_world = 0


def native_allreduce_init(rank: int, world: int) -> None:
    """Communicator of the library's own NCCL entry points (``scann_allreduce_*``, include/scann_b200.h): rank 0 creates
    the unique id, torch.distributed carries its 128 bytes to the other ranks (any host channel would do), every rank
    joins with its CUDA device current.  One communicator per process."""
    global _native_world
    import ctypes as C
    from ._abi import check, lib
    if _native_world:
        if _native_world != world:
            raise RuntimeError("the native NCCL communicator was created for another world size")
        return
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib.scann_allreduce_unique_id(C.cast(buf, C.c_void_p)), "allreduce_unique_id")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0)
    buf = (C.c_ubyte * 128).from_buffer_copy(box[0])
    check(lib.scann_allreduce_init(C.cast(buf, C.c_void_p), rank, world), "allreduce_init")
    _native_world = world


def native_allreduce_sum(t: torch.Tensor) -> None:
    """In-place fp32 sum over the ranks through ``scann_allreduce_sum`` on the current stream (capturable)."""
    from ._abi import check, lib
    assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    check(lib.scann_allreduce_sum(t.data_ptr(), t.numel(), torch.cuda.current_stream(t.device).cuda_stream), "allreduce_sum")


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n structures for this rank (SURVEY.md 8e partitioning)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def balanced_shards(cost, world: int):
    """Assign len(cost) structures to ``world`` ranks, the same NUMBER of structures per rank (len(cost) must be a
    multiple of world -- the RMSE of the global batch does not care which rank holds which structure) and nearly the
    same total ``cost`` (valid atom-neighbour pairs: what the local-attention kernels' time is proportional to).
    A data-parallel step ends with an exchange every rank waits for, so its time is the SLOWEST rank's: shards of
    randomly drawn structures differ by 3 % (QM9) to 11 % (MP2018-shaped crystals) in their pair counts, and that
    difference is lost on every step.  Longest-processing-time greedy: structures in decreasing cost order, each to
    the lightest rank that still has room.  Returns a list of ``world`` sorted index arrays."""
    import numpy as np
    cost = np.asarray(cost, np.int64)
    n = len(cost)
    if n % world:
        raise ValueError("the number of structures must be a multiple of the number of ranks")
    per = n // world
    load = np.zeros(world, np.int64)
    count = np.zeros(world, np.int64)
    owner = np.empty(n, np.int64)
    for i in np.argsort(-cost, kind="stable"):
        open_ranks = np.flatnonzero(count < per)
        r = open_ranks[np.argmin(load[open_ranks])]
        owner[i] = r
        load[r] += cost[i]
        count[r] += 1
    return [np.flatnonzero(owner == r) for r in range(world)]


class P2PExchange:
    """Gradient exchange fused into the optimiser kernel over NVLink peer memory (scann_b200/csrc/p2p.cu): every rank's
    gradient arena sits in a block the other ranks of the node map through CUDA IPC; ``scann_adam_p2p_step`` sums the
    peers' arenas while it updates.  Replaces ``allreduce(grads)`` + ``scann_adam_step`` of the NCCL path; the engine's
    ``grads`` tensor is re-pointed at the shared block, so call this before the first train step."""

    MAX_RANKS = 8

    def __init__(self, engine, rank: int, world: int):
        import ctypes as C
        import numpy as np
        from ._abi import check, lib
        if world > self.MAX_RANKS:
            raise ValueError(f"peer-memory exchange supports up to {self.MAX_RANKS} ranks of one node")
        n = engine.layout.total
        nfl = (n + 4 + 63) // 64 * 64
        nbytes = nfl * 4 + 64 * 4 + nfl * 4       # arena | 64 flag words | reduced arena (two-shot exchange, >= 4 ranks)
        ptr = C.c_void_p()
        check(lib.scann_p2p_alloc(nbytes, C.byref(ptr)), "p2p_alloc")
        self.local = int(ptr.value)
        handle = (C.c_ubyte * 64)()
        check(lib.scann_p2p_export(self.local, C.cast(handle, C.c_void_p)), "p2p_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        self.bases = []
        for r in range(world):
            if r == rank:
                self.bases.append(self.local)
                continue
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
            peer = C.c_void_p()
            check(lib.scann_p2p_import(C.cast(buf, C.c_void_p), C.byref(peer)), f"p2p_import(rank {r})")
            self.bases.append(int(peer.value))
        # ScannP2PBlock: const float* arena[8]; uint32_t* flags[8]; int world, rank
        blk = np.zeros(17, np.int64)
        for r, base in enumerate(self.bases):
            blk[r] = base
            blk[8 + r] = base + nfl * 4
        blk[16] = (rank << 32) | world          # little endian: {int world, int rank}
        dev = engine.device
        self.block = torch.from_numpy(blk).to(dev)
        self.sums = torch.zeros(4, dtype=torch.float32, device=dev)

        class _Raw:                              # zero-copy torch view of the shared block's arena
            __cuda_array_interface__ = {"shape": (n + 4,), "typestr": "<f4", "data": (self.local, False), "version": 3}

        self._raw = _Raw()
        engine.grads = torch.as_tensor(self._raw, device=dev)
        assert engine.grads.data_ptr() == self.local
        engine.p2p = self
        self.rank, self.world = rank, world
        dist.barrier()                           # every rank has mapped every block before anyone starts a step


def attach(model, world: int) -> None:
    """Make ``model.train_on_batch`` data-parallel: local shard in, global-batch semantics out.  The exchange is the
    peer-memory form fused into the optimiser kernel (``P2PExchange``; ranks of one node, the default there) or one
    NCCL all-reduce of the gradient arena per step (SCANN_P2P_REDUCE=0, or ranks on several nodes)."""
    if world > 1:
        model.allreduce = allreduce_sum
        model.world_size = world
        # SCANN_NCCL=native: the all-reduce goes through the library's own scann_allreduce_* entry points (NCCL bound at
        # run time) instead of torch.distributed's all_reduce -- same collective, same place in the captured graph
        if os.environ.get("SCANN_NCCL", "torch") == "native" and torch.cuda.is_available():
            native_allreduce_init(env_world()[0], world)
            model.allreduce = native_allreduce_sum
        # default on one node: the peer-memory exchange (validated against the single-GPU full-batch run on 2, 4 and 8
        # GPUs, tools/dp_check.py); SCANN_P2P_REDUCE=0 selects the NCCL all-reduce, which is also the multi-node path
        one_node = int(os.environ.get("LOCAL_WORLD_SIZE", str(world))) == world and world <= P2PExchange.MAX_RANKS
        want = os.environ.get("SCANN_P2P_REDUCE", "1" if one_node else "0") == "1"
        if want and torch.cuda.is_available():
            P2PExchange(model.engine, env_world()[0], world)
            model.allreduce = None
        # every rank draws its own Dropout masks (its shard holds different structures)
        rank = env_world()[0]
        model.engine.dropout_seed = (model.engine.dropout_seed + 7919 * rank) & 0x7FFFFFFF
