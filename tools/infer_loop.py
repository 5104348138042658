"""A few inference forwards (and optionally train steps) without CUDA graphs: the command line ncu profiles."""
import os, sys, torch
os.environ.setdefault("SCANN_GRAPHS", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
train = len(sys.argv) > 3 and sys.argv[3] == "train"
m = create_model(get_config("qm9")); eng = m.engine
inp, tgt = make_batch("qm9", 0, B=B)
b = eng.load_batch(inp, plan=False)
t = torch.from_numpy(tgt).cuda()
for _ in range(n):
    if train:
        eng.train_step(b, t, 5e-4, replan=True)
    else:
        eng.predict_step(b, replan=True)
torch.cuda.synchronize()
eng.check_status()
print("ok", eng.launches)
