#!/bin/bash
# error of y / ga_score / gradients at full depth and batch against the fp64 oracle, for the current build
O=gpurun_out; TAG=${1:-r02num}
python - > $O/${TAG}.log 2>&1 <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from oracle import scann_oracle as O
from scann_b200.config import model_spec
from scann_b200.configs import get_config
from scann_b200.params import ParamLayout
from scann_b200.synth import make_batch
from scann_b200.engine import Engine
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())
for name, B in (("qm9", 128), ("mp2018", 64), ("fullerene", 128)):
    cfg = get_config(name); spec = model_spec(cfg); lay = ParamLayout(spec); arena = lay.randomize_arena(21)
    inputs, target = make_batch(name, 13, B=B)
    kw = dict(n_attention=spec.n_attention, g_update=spec.g_update, gaussian_d=spec.gaussian_d, use_attn_norm=spec.use_attn_norm, use_ga_norm=spec.use_ga_norm)
    w = lay.to_dict(arena)
    y64, ga64 = O.predict(w, inputs, torch.float64, **kw)
    y32, ga32 = O.predict(w, inputs, torch.float32, **kw)
    for pipe in ("15", "0"):
        os.environ["SCANN_LA_PIPE"] = pipe
        eng = Engine(spec, arena); b = eng.load_batch(inputs); y, ga = eng.forward(b); torch.cuda.synchronize(); eng.check_status()
        ga = ga.cpu().numpy().reshape(b.B, b.M)
        e = np.abs(ga - ga64[..., 0]).max(1) / np.abs(ga64).max()
        print(f"{name} B={B} pipe={pipe} stride={b.stride}: y gpu {rel(y.cpu().numpy(), y64.ravel()):.2e} (fp32 oracle {rel(y32, y64):.2e})  ga gpu {rel(ga, ga64[..., 0]):.2e} (fp32 oracle {rel(ga32, ga64):.2e})  worst structures {np.argsort(e)[-3:].tolist()} atoms {inputs['atom_mask'][np.argsort(e)[-3:]].sum((1,2)).tolist()}")
PY
cat $O/${TAG}.log
