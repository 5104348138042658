"""A few inference forwards / train steps of one workload without CUDA graphs: the command line ncu profiles.
usage: wl_loop.py workload steps [train]"""
import os, sys, torch
os.environ.setdefault("SCANN_GRAPHS", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch
wl = sys.argv[1] if len(sys.argv) > 1 else "qm9"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
train = len(sys.argv) > 3 and sys.argv[3] == "train"
cfg = get_config(wl)
if wl == "ptgp":      # model_ptgp.yaml lacks these two keys (KeyError in the reference as shipped)
    cfg["model"].update(g_update=False, gaussian_d=4.0)
m = create_model(cfg); eng = m.engine
inp, tgt = make_batch(wl, 0, use_ring=bool(cfg["model"].get("use_ring")))
b = eng.load_batch(inp, plan=False)
t = torch.from_numpy(tgt).cuda()
for _ in range(n):
    if train:
        eng.train_step(b, t, 5e-4, replan=True)
    else:
        eng.predict_step(b, replan=True)
torch.cuda.synchronize()
eng.check_status()
print("ok", eng.launches, "stride", b.stride, "rows", b.rows)
