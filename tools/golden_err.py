"""Errors of the engine against the committed goldens (fp64 oracle), as numbers instead of pass / fail."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden.make_golden import CASES, build_case
from scann_b200.engine import Engine
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
for name in CASES:
    cfg, spec, lay, arena, inputs, target = build_case(name)
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"{name}.npz"))
    eng = Engine(spec, arena)
    b = eng.load_batch(inputs)
    y, ga = eng.forward(b)
    torch.cuda.synchronize(); eng.check_status()
    ey, ega = rel(y.cpu().numpy(), z["y"].ravel()), rel(ga.cpu().numpy().reshape(b.B, b.M), z["ga"][..., 0])
    eng.train_step(b, torch.from_numpy(target).cuda(), lr=1e-3, apply=False, want_grads=True)
    torch.cuda.synchronize(); eng.check_status()
    g = eng.grad_out.cpu().numpy().astype(np.float64)
    print(f"{name}: y {ey:.2e}  ga {ega:.2e}  grad sample {rel(g[z['grad_idx']], z['grad_sample']):.2e}")
