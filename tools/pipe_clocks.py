"""Phase timestamps of the pipelined geometry kernel (CTA 0, consumer group 0): build with SCANN_NVCC_DEFS=-DSCANN_PIPE_CLK."""
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
for _ in range(3):
    eng.predict_step(b, replan=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 96)()
assert raw.scann_pipe_clocks(buf) == 0
a = np.array(list(buf), np.int64).reshape(2, 4, 12)
order = [0, 1, 9, 2, 3, 4, 5, 6, 7, 8]
names = {0: "loop top", 1: "tile landed", 9: "gathers issued, split, fence", 2: "group sync + MMA chains issued", 3: "accumulators ready",
         4: "gathered rows arrived", 5: "acc -> image + sync", 6: "epilogue", 7: "fence + sync", 8: "handed to the store warp"}
g = a[0]
t00 = g[0][10]
print(f"B={B} ntiles={int(b.ntiles.item())}  (cycles since the kernel's tensor-memory allocation)")
print(f"weights -> tensor memory: {g[0][11] - g[0][10]} cycles")
for k in range(4):
    if g[k][1] == 0: continue
    prev = g[k][0]
    parts = []
    for p in order:
        parts.append(f"{names[p]} +{g[k][p] - prev}")
        prev = g[k][p]
    print(f"tile ordinal {k} (@{g[k][0] - t00}, total {g[k][8] - g[k][0]}): " + ", ".join(parts))

names1 = ["loop top", "tile landed", "a = x[j] * g', lo, fence, sync", "MMA issued", "accumulators ready", "acc -> image (keys) + sync",
          "scores + sync", "per-atom softmax / context / LayerNorm", "sync"]
g = a[1]
print("--- attention forward")
for k in range(4):
    if g[k][1] == 0: continue
    print(f"tile ordinal {k} (total {g[k][8] - g[k][0]}): " + ", ".join(f"{names1[p]} +{g[k][p] - g[k][p-1]}" for p in range(1, 9)))
