"""In-graph cost of kernel groups: train-step time with a group left out (SCANN_DEBUG_SKIP) vs the full step."""
import os, subprocess, sys, json
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import os, sys, torch
sys.path.insert(0, os.path.dirname(%r))
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch
m = create_model(get_config("qm9")); eng = m.engine
inp, tgt = make_batch("qm9", 0, B=128)
b = eng.load_batch(inp, plan=False)
t = torch.from_numpy(tgt).cuda()
for _ in range(4): eng.train_step(b, t, 5e-4, replan=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): eng.train_step(b, t, 5e-4, replan=True)
e1.record(); torch.cuda.synchronize()
print(e0.elapsed_time(e1) / 30)
''' % here
base = None
for skip in ["", "wgrad", "la_fwd", "la_bwd", "la_fwd,la_bwd,wgrad", "rn", "geom_init", "chain", "la_fwd,la_bwd,wgrad,chain,geom_init"]:
    env = dict(os.environ, SCANN_DEBUG_SKIP=skip)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    try:
        ms = float(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(skip, "FAILED", out.stderr[-300:]); continue
    base = base or ms
    print(f"skip={skip or '-':24s} {ms:.3f} ms  (delta {base - ms:+.3f})")
