"""World-size-2 gloo test of the data-parallel semantics (host logic, CPU only).

The reference loss sqrt(mean_B err^2) is not additive over shards (losses.py:5-6).  The engine
therefore back-propagates G = sum_b err_b dy_b/dtheta per rank, all-reduces [G, SSE] ONCE and
applies 1/(B*RMSE) + the l2 gradient afterwards.  Here the per-rank part is played by the
oracle; the reduced result must equal the single-process full-batch gradient of the reference loss.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import scann_oracle as O
from scann_b200 import dist as sdist
from scann_b200.config import model_spec
from scann_b200.configs import get_config
from scann_b200.params import ParamLayout
from scann_b200.synth import make_batch


def _case():
    cfg = get_config("qm9")
    cfg["model"]["n_attention"] = 2
    spec = model_spec(cfg)
    lay = ParamLayout(spec)
    arena = lay.randomize_arena(7)
    inputs, target = make_batch("qm9", 21, B=6)
    kw = dict(n_attention=2, g_update=True, gaussian_d=spec.gaussian_d, use_attn_norm=True, use_ga_norm=True)
    return lay, arena, inputs, target, kw


def _shard_G(lay, arena, inputs, target, kw, lo, hi):
    w = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in lay.to_dict(arena).items()}
    sh = {k: v[lo:hi] for k, v in inputs.items()}
    y, _ = O.forward(w, O.to_torch_inputs(sh), **kw)
    err = (y.reshape(-1) - torch.tensor(target[lo:hi], dtype=torch.float64)).detach()
    (err * y.reshape(-1)).sum().backward()
    g = np.zeros(lay.total)                      # float64 arena (ParamLayout.from_dict packs float32)
    for e in lay:
        if w[e.name].grad is not None:
            g[e.offset:e.offset + e.size] = w[e.name].grad.numpy().reshape(-1)
    return g, float((err ** 2).sum())


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    r, _, w_ = sdist.init("gloo")
    lay, arena, inputs, target, kw = _case()
    lo, hi = sdist.shard_bounds(len(target), r, w_)
    G, sse = _shard_G(lay, arena, inputs, target, kw, lo, hi)
    buf = torch.tensor(np.concatenate([G, [sse]]))
    sdist.allreduce_sum(buf)
    if rank == 0:
        q.put(buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_gradient_equals_full_batch_gradient():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    red = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    lay, arena, inputs, target, kw = _case()
    B = len(target)
    G, sse = red[:-1], red[-1]
    rmse = np.sqrt(sse / B)
    grad_dp = G / (B * rmse) + 2 * O.L2_COEF * lay.l2_mask() * arena          # what scann_adam_step applies
    l2n = [e.name for e in lay if e.l2]
    loss, _, _, grads = O.loss_and_grads(lay.to_dict(arena), inputs, target, l2n, **kw)
    ref = np.zeros(lay.total)
    for e in lay:
        ref[e.offset:e.offset + e.size] = grads[e.name].reshape(-1)
    assert abs(rmse + O.L2_COEF * float((lay.l2_mask() * arena.astype(np.float64) ** 2).sum()) - loss) < 1e-9
    np.testing.assert_allclose(grad_dp, ref, rtol=1e-7, atol=1e-10)


def test_shard_bounds_cover_batch():
    for n in (1, 7, 64, 129):
        for world in (1, 2, 4, 8):
            spans = [sdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
