"""Phase timestamps of the pipelined backward kernels (CTA 0, consumer group 0, last launch = layer 0): -DSCANN_DEV_PROBES build."""
import ctypes, os, sys, numpy as np, torch
os.environ.setdefault("SCANN_GRAPHS", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200 import _abi
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch
raw = ctypes.CDLL(_abi.LIB_PATH)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = create_model(get_config("qm9")); eng = m.engine
inp, tgt = make_batch("qm9", 0, B=B)
b = eng.load_batch(inp, plan=False)
t = torch.from_numpy(tgt).cuda()
for _ in range(3):
    eng.train_step(b, t, 5e-4, replan=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 96)()
assert raw.scann_pipe_clocks_bwd(buf) == 0
a = np.array(list(buf), np.int64).reshape(2, 4, 12)
names = [["loop top", "tile landed", "phase A (scores) + sync", "B1 (softmax bwd per head) + sync", "B2 (dq per atom) + sync",
          "C (d_k, lo) + fence + sync", "MMA issued", "accumulators ready (+ next rows issued)", "acc -> image + sync", "D (scatter, d_a x) + fence + sync"],
         ["loop top", "tile landed", "A (LN_g bwd, d_pre, scatter) + fence + sync", "MMA issued", "B (s_pre per atom)", "accumulators ready",
          "acc -> image + sync", "C (dg) + fence + sync"]]
for k, kname in enumerate(("attention backward", "geometry backward")):
    print(f"--- {kname}, B={B} ntiles={int(b.ntiles.item())}")
    for o in range(4):
        g = a[k][o]
        if g[1] == 0: continue
        n = len(names[k])
        print(f"tile ordinal {o} (total {g[n-1] - g[0]}): " + ", ".join(f"{names[k][p]} +{g[p] - g[p-1]}" for p in range(1, n)))
