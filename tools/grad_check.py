"""Per-tensor gradient error of a golden case against the fp64 oracle (development tool, GPU only)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scann_oracle as O
from scann_b200.engine import Engine
from tests.golden.make_golden import build_case, oracle_kwargs

name = sys.argv[1] if len(sys.argv) > 1 else "fullerene_b2_l2"
cfg, spec, lay, arena, inputs, target = build_case(name)
w = lay.to_dict(arena); l2n = [e.name for e in lay if e.l2]
loss, y, ga, g64 = O.loss_and_grads(w, inputs, target, l2n, **oracle_kwargs(spec))
eng = Engine(spec, arena)
b = eng.load_batch(inputs)
eng.train_step(b, torch.from_numpy(target).cuda(), lr=1e-3, apply=False, want_grads=True)
torch.cuda.synchronize(); eng.check_status()
g = lay.to_dict(eng.grad_out.cpu().numpy())
gmax = max(np.abs(v).max() for v in g64.values())
rows = []
for e in lay:
    err = np.abs(g[e.name].astype(np.float64) - g64[e.name]).max()
    rows.append((err / gmax, err / max(np.abs(g64[e.name]).max(), 1e-30), e.name, np.abs(g64[e.name]).max()))
rows.sort(reverse=True)
print("case", name, "pairs", int(inputs["neighbor_mask"].sum()), "gmax", gmax)
for r in rows[:12]:
    print(f"{r[2]:45s} err/gmax {r[0]:.2e}  err/|t|max {r[1]:.2e}  |t|max {r[3]:.2e}")
