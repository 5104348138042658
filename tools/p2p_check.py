"""Peer-memory gradient exchange (SCANN_P2P_REDUCE) against the NCCL all-reduce path: same parameters after a few
data-parallel train steps, bit-identical parameters on all ranks.  Run under torchrun with >= 2 ranks."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200 import dist as sdist
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch

rank, local_rank, world = sdist.init()
torch.cuda.set_device(local_rank)
cfg = get_config("qm9"); cfg["model"]["n_attention"] = 2


def run(p2p: bool):
    os.environ["SCANN_P2P_REDUCE"] = "1" if p2p else "0"
    m = create_model(cfg, seed=3)
    sdist.attach(m, world)
    m.dropout = False
    losses = []
    for step in range(4):
        # same masks on every rank (ranks must meet the same batch shapes in the same order: a new shape's first,
        # eager pass of the NCCL path performs an all-reduce), different geometry and targets
        inp, tgt = make_batch("qm9", 10 * step, B=16)
        inp["neighbor_distance"] = (inp["neighbor_distance"] * (1.0 + 0.05 * rank)).astype(np.float32)
        inp["neighbor_weight"] = (inp["neighbor_weight"] * (1.0 - 0.03 * rank)).astype(np.float32)
        tgt = (tgt + 0.1 * rank).astype(np.float32)
        losses.append(m.train_on_batch(inp, tgt))
    return m.engine.get_params(), losses


pa, la = run(False)
pb, lb = run(True)
err = float(np.abs(pa - pb).max() / np.abs(pa).max())
mine = torch.from_numpy(pb).cuda()
allp = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allp, mine)
same = all(bool(torch.equal(allp[0], t)) for t in allp)
if rank == 0:
    print(f"world {world}: params nccl vs p2p rel err {err:.2e}; losses nccl {la} p2p {lb}; identical across ranks: {same}")
    assert err <= 1e-5 and same and np.allclose(la, lb, rtol=1e-5)
    print("P2P_CHECK_OK")
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
