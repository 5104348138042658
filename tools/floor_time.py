"""Train-step time vs batch size in CUDA-graph mode: how much of the step is fixed latency."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch, count_valid
m = create_model(get_config("qm9")); eng = m.engine
for B in (2, 16, 64, 128, 256, 512):
    inp, tgt = make_batch("qm9", 0, B=B)
    b = eng.load_batch(inp, plan=False)
    t = torch.from_numpy(tgt).cuda()
    for _ in range(4): eng.train_step(b, t, 5e-4, replan=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.train_step(b, t, 5e-4, replan=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    for _ in range(4): eng.predict_step(b, replan=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): eng.predict_step(b, replan=True)
    e1.record(); torch.cuda.synchronize()
    msf = e0.elapsed_time(e1) / 20
    print(f"B={B:4d} pairs={count_valid(inp)[1]:6d} tiles={int(b.ntiles.item()):4d} train {ms:.3f} ms ({B/ms*1e3:8.0f}/s)  infer {msf:.3f} ms ({B/msf*1e3:8.0f}/s)")
