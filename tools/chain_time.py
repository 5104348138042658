"""Back-to-back launches of ONE chained-Dense launch (CUDA events): the step lists of a real train step are recorded
and replayed through scann_dense_chain (round 1) and scann_dense_chain2 (warp-specialised), for every prefix length.
With a -DSCANN_DEV_PROBES build the phase clocks of CTA 0 of the last chain2 launch are printed as well.
usage: chain_time.py [shape] [B]"""
import ctypes, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SCANN_GRAPHS"] = "0"
shape = sys.argv[1] if len(sys.argv) > 1 else "qm9"
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
from scann_b200._abi import lib, check, ChainStep
from scann_b200.configs import get_config
from scann_b200.engine import Engine
from scann_b200.config import model_spec
from scann_b200.params import ParamLayout
from scann_b200.synth import make_batch

cfg = get_config(shape); spec = model_spec(cfg); lay = ParamLayout(spec)
eng = Engine(spec, lay.randomize_arena(1))
inp, tgt = make_batch(shape, 0, B=B)
b = eng.load_batch(inp)
t = torch.from_numpy(tgt).cuda()
rec = []
orig = eng._chain
def recorder(steps, R):
    rec.append(([ChainStep.from_buffer_copy(bytes(s)) for s in steps], R))
    orig(steps, R)
eng._chain = recorder
eng.train_step(b, t, 1e-3, apply=False)
torch.cuda.synchronize(); eng.check_status()
eng._chain = orig
st = eng._stream()
L = spec.n_attention

def launch(steps, R, new):
    steps = [ChainStep.from_buffer_copy(bytes(s)) for s in steps]
    if new:
        for s in steps:
            for kb in range(s.kblk):
                s.W[kb] = eng._wimg_ptr(s.W[kb])
    arr = (ChainStep * len(steps))(*steps)
    fn = lib.scann_dense_chain2 if new else lib.scann_dense_chain
    check(fn(ctypes.cast(arr, ctypes.c_void_p), len(steps), R, 0, st) if new else fn(ctypes.cast(arr, ctypes.c_void_p), len(steps), R, st))

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

names = {1: "forward: ResidualNorm_0 + projections of layer 1", L + 1: "backward head: GA^T, after_Lc^T, tail",
         L + 2: "backward: x-gradient of a layer + tail of the layer below"}
for idx, name in names.items():
    steps, R = rec[idx]
    print(f"--- {shape} R={R} {name}: {len(steps)} steps, kblk {[s.kblk for s in steps]}")
    for n in range(1, len(steps) + 1):
        row = []
        for pdl in (0, 1):
            lib.scann_set_pdl(pdl)
            row.append((timeit(lambda: launch(steps[:n], R, False)), timeit(lambda: launch(steps[:n], R, True))))
        lib.scann_set_pdl(0)
        print(f"  first {n} steps ({sum(s.kblk for s in steps[:n])} weight blocks): round-1 {row[0][0]:.1f} us, chain2 {row[0][1]:.1f} us"
              f" | with PDL: {row[1][0]:.1f} / {row[1][1]:.1f} us")
    if hasattr(lib, "scann_debug_clocks_chain2"):
        launch(steps, R, True); torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 192)()
        check(lib.scann_debug_clocks_chain2(ctypes.cast(buf, ctypes.c_void_p)))
        c = np.array(buf[:], dtype=np.int64).reshape(3, 64)
        t0 = c[0, 0]
        print("  epilogue thread 0 (cycles after its PDL wait): " + " | ".join(
            f"step {si}: img {c[0,1+4*si]-t0} acc {c[0,2+4*si]-t0} readback {c[0,3+4*si]-t0} done {c[0,4+4*si]-t0}" for si in range(len(steps))))
        print("  row epilogue of thread 0 (relative to its start): " + " | ".join(
            f"step {si} mode {steps[si].mode}: v {c[2,6*si+1]-c[2,6*si]} mode {c[2,6*si+2]-c[2,6*si]} stores {c[2,6*si+3]-c[2,6*si]} "
            f"cnt {c[2,6*si+4]-c[2,6*si]} red {c[2,6*si+5]-c[2,6*si]} (start {c[2,6*si]-t0})" for si in range(len(steps))))
        nb = sum(s.kblk for s in steps)
        print("  MMA thread: " + " | ".join(f"blk {bi}: go {c[1,3*bi]-t0} lastslot {c[1,3*bi+1]-t0} committed {c[1,3*bi+2]-t0}" for bi in range(nb)))
eng.check_status()
