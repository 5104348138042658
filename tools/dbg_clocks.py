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
buf = (ctypes.c_longlong * 32)()
check(lib.scann_debug_clocks(ctypes.cast(buf, ctypes.c_void_p)))
c = np.array(buf[:9], dtype=np.int64)
names = ["weights->TMEM", "stage tile", "issue MMA", "prefetch gathers", "wait MMA", "TMEM->smem", "row epilogue", "rest (2nd tile etc)"]
for n, d in zip(names, np.diff(c)): print(f"{n:22s} {d:8d} cycles")
print("total", c[8] - c[0])

buf2 = (ctypes.c_longlong * 16)()
check(lib.scann_debug_clocks_dense(ctypes.cast(buf2, ctypes.c_void_p)))
c = np.array(buf2[:7], dtype=np.int64)
print("dealloc", c[6]-c[4], "epilogue2 after dealloc", c[5]-c[6])
c = c[:6]
print("dense_tc (last launch = GA q/k projection, kblk=1):")
for n, d in zip(["alloc+init", "W->TMEM + stage X", "MMA issue+wait", "TMEM->smem", "row epilogue"], np.diff(c)): print(f"{n:22s} {d:8d} cycles")
print("total", c[5] - c[0])
