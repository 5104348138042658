"""A/B timing of engine switches: train-step and inference time (CUDA events, graphs on) per environment setting.
usage: python tools/ab_time.py [workload[:B]] VAR=val[,VAR=val] ...     ('-' = defaults)"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import os, sys, torch
sys.path.insert(0, os.path.dirname(%r))
from scann_b200.configs import get_config
from scann_b200.model import create_model
from scann_b200.synth import make_batch
wl, B = sys.argv[1], int(sys.argv[2])
cfg = get_config(wl)
if wl == "ptgp":      # model_ptgp.yaml lacks these two keys (KeyError in the reference as shipped)
    cfg["model"].update(g_update=False, gaussian_d=4.0)
m = create_model(cfg); eng = m.engine
ring = bool(cfg["model"].get("use_ring"))
inp, tgt = make_batch(wl, 0, B=B, use_ring=ring) if B else make_batch(wl, 0, use_ring=ring)
b = eng.load_batch(inp, plan=False)
t = torch.from_numpy(tgt).cuda()
for _ in range(4): eng.train_step(b, t, 5e-4, replan=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): eng.train_step(b, t, 5e-4, replan=True)
e1.record(); torch.cuda.synchronize()
tr = e0.elapsed_time(e1) / 30
for _ in range(4): eng.predict_step(b, replan=True)
torch.cuda.synchronize()
e0.record()
for _ in range(30): eng.predict_step(b, replan=True)
e1.record(); torch.cuda.synchronize()
eng.check_status()
print(tr, e0.elapsed_time(e1) / 30)
''' % here
args = sys.argv[1:]
wl, B = "qm9", 0
if args and "=" not in args[0] and args[0] != "-":
    wl = args.pop(0)
    if ":" in wl:
        wl, B = wl.split(":")[0], int(wl.split(":")[1])
for setting in args or ["-"]:
    env = dict(os.environ)
    if setting != "-":
        for kv in setting.split(","):
            k, v = kv.split("=", 1)
            env[k] = v
    out = subprocess.run([sys.executable, "-c", code, wl, str(B)], env=env, capture_output=True, text=True)
    try:
        tr, inf = map(float, out.stdout.strip().splitlines()[-1].split())
        print(f"{wl}:{B or 'default'} {setting:40s} train {tr:.3f} ms  infer {inf:.3f} ms")
    except Exception:
        print(setting, "FAILED", out.stdout[-300:], out.stderr[-600:])
