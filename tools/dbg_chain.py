"""Phase times (cycles) of CTA 0 of the last dense_chain launch of an inference forward (= the final chain)."""
import os, sys, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200._abi import lib, check
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch
m = create_model(get_config("qm9")); eng = m.engine; eng.use_graphs = False
inp, tgt = make_batch("qm9", 0, B=128)
b = eng.load_batch(inp)
for _ in range(3): eng.forward(b)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
check(lib.scann_debug_clocks_chain(ctypes.cast(buf, ctypes.c_void_p)))
c = np.array(buf[:], dtype=np.int64)
print("prologue: W0->TMEM", c[1] - c[0], " pdl_wait", c[2] - c[1])
prev = c[2]
for si in range(5):
    t = c[3 + si * 4: 7 + si * 4]
    print(f"step {si}: stage(W+X) {t[0] - prev:6d}  mma {t[1] - t[0]:6d}  tmem->S {t[2] - t[1]:6d}  epilogue {t[3] - t[2]:6d}")
    prev = t[3]
print("total", prev - c[0])
