"""Data-parallel train steps on >= 2 GPUs against the single-GPU full-batch run (the reference's semantics: one
device, the whole batch, loss = sqrt(mean_B err^2), scann/layers/losses.py:5-6).

Run under torchrun with WORLD_SIZE ranks.  Every step, rank r trains on its shard of a 2*8*WORLD_SIZE... batch whose
shards have DIFFERENT pair counts (so the ranks' shape-class / CUDA-graph caches miss at different steps), once with
torch.distributed's NCCL all-reduce, once with the library's own scann_allreduce_* entry points (SCANN_NCCL=native) and once
with the peer-memory exchange (SCANN_P2P_REDUCE=1); rank 0 also trains a single-GPU model
on the concatenated batches.  Parameters after the steps must agree and be bit-identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200 import dist as sdist                     # noqa: E402
from scann_b200.configs import get_config                # noqa: E402
from scann_b200.model import create_model                # noqa: E402
from scann_b200.synth import make_batch                  # noqa: E402

STEPS = int(os.environ.get("DP_CHECK_STEPS", "3"))
PER = 8
rank, local_rank, world = sdist.init()
torch.cuda.set_device(local_rank)
cfg = get_config("qm9")
cfg["model"]["n_attention"] = 2


def batches():
    out = []
    for step in range(STEPS):
        # step 1 repeats step 0's shapes on rank 0 only: the ranks' graph caches hit / miss at different steps
        inp, tgt = make_batch("qm9", 100 + (step if step != 1 else 0), B=PER * world)
        if step == 1:
            inp2, tgt2 = make_batch("qm9", 777, B=PER * world)
            for k in inp:
                inp[k] = np.concatenate([inp[k][:PER], inp2[k][PER:]])
            tgt = np.concatenate([tgt[:PER], tgt2[PER:]])
        out.append((inp, tgt))
    return out


def run(mode: str):
    os.environ["SCANN_P2P_REDUCE"] = "1" if mode == "p2p" else "0"
    os.environ["SCANN_NCCL"] = "native" if mode == "nccl-native" else "torch"
    m = create_model(cfg, seed=3)
    m.dropout = False
    if mode != "single":
        sdist.attach(m, world)
    losses = []
    for inp, tgt in batches():
        if mode != "single":
            lo, hi = rank * PER, (rank + 1) * PER
            inp, tgt = {k: v[lo:hi] for k, v in inp.items()}, tgt[lo:hi]
        losses.append(m.train_on_batch(inp, tgt))
    return m.engine.get_params().astype(np.float64), losses, m.engine.layout.init_arena(3).astype(np.float64)


ok = True
ref = None
if rank == 0:
    ref, lref, p0 = run("single")
for mode in ("nccl", "nccl-native", "p2p"):
    dist.barrier()
    p, losses, p0 = run(mode)
    mine = torch.from_numpy(p).cuda()
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    same = all(bool(torch.equal(allp[0], t)) for t in allp)
    if rank == 0:
        moved = np.abs(ref - p0).max()
        err = float(np.abs(p - ref).max() / moved)
        lerr = float(np.abs(np.array(losses) - np.array(lref)).max() / np.abs(lref).max())
        print(f"world {world} {mode}: parameter movement error vs single-GPU full batch {err:.2e}, loss error {lerr:.2e}, "
              f"identical across ranks: {same}")
        ok = ok and err <= 2e-2 and lerr <= 1e-5 and same
if rank == 0:
    print("DP_CHECK_OK" if ok else "DP_CHECK_FAILED")
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0 if ok else 1)
