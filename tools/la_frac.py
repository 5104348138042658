"""Roofline fractions of the local-attention kernels at several batch sizes (CUDA events around the launches of one layer,
eager launches behind a spin kernel, L2 flushed between steps -- the method of bench.py's `roofline` entry)."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch, count_valid
peak = 6540.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
shape = sys.argv[1] if len(sys.argv) > 1 else "qm9"
m = create_model(get_config(shape)); eng = m.engine
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for B in [int(x) for x in (sys.argv[2:] or ["128", "512"])]:
    inp, tgt = make_batch(shape, 0, B=B)
    A, P = count_valid(inp)
    b = eng.load_batch(inp, plan=False)
    t = torch.from_numpy(tgt).cuda()
    for _ in range(3): eng.train_step(b, t, 5e-4, replan=True)
    torch.cuda.synchronize()
    eng.prof = {}
    for _ in range(10):
        flush.fill_(1.0)
        torch.cuda._sleep(6_000_000)
        eng.train_step(b, t, 5e-4, replan=True)
    torch.cuda.synchronize()
    prof = eng.prof_summary(); eng.prof = None
    nf, msf = prof["la_forward"]; nb, msb = prof["la_backward"]; nw, msw = prof.get("wgrad_batch", (0, float("nan")))
    bf, bb = 1028.0 * P + 1028.0 * A, 1540.0 * P + 1028.0 * A
    L = m.engine.spec.n_attention
    bw = L * (4 * 512.0 * P + 9 * 512.0 * A)
    print(f"{shape} B={B} pairs={P} atoms={A} tiles={int(b.ntiles.item())}: LA forward {msf*1e3:.1f} us / layer = {bf/msf/1e6:.0f} GB/s = {bf/msf/1e6/peak:.3f} of {peak:.0f} | "
          f"LA backward {msb*1e3:.1f} us = {bb/msb/1e6:.0f} GB/s = {bb/msb/1e6/peak:.3f} | weight gradients {msw*1e3:.0f} us = {bw/msw/1e6/peak:.3f}")
