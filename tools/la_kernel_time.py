"""Back-to-back launches of ONE local-attention kernel on warm buffers (CUDA events): microseconds per launch for the
pipelined and the round-1 forms, forward and backward halves.  usage: la_kernel_time.py [B] [shape]"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
shape = sys.argv[2] if len(sys.argv) > 2 else "qm9"
from scann_b200._abi import lib, check
from scann_b200.configs import get_config
from scann_b200.engine import Engine, _p, D
from scann_b200.config import model_spec
from scann_b200.params import ParamLayout, layer_name
from scann_b200.synth import make_batch

def run(pipe, stride):
    os.environ["SCANN_LA_PIPE"] = str(pipe)
    os.environ["SCANN_TILE_STRIDE"] = str(stride)
    os.environ["SCANN_GRAPHS"] = "0"
    cfg = get_config(shape); spec = model_spec(cfg); lay = ParamLayout(spec)
    eng = Engine(spec, lay.randomize_arena(1))
    inp, tgt = make_batch(shape, 0, B=B)
    b = eng.load_batch(inp)
    t = torch.from_numpy(tgt).cuda()
    eng.train_step(b, t, 1e-3, apply=False)          # fills every workspace buffer of a training step
    torch.cuda.synchronize(); eng.check_status()
    ws = eng._workspace(b, True)
    st = eng._stream()
    l = 1
    la = layer_name("local_attention", l); fg = f"{la}/filter_geo/kernel"
    la_args = (_p(b.ntiles), _p(b.tile_a0), _p(b.tile_a1), _p(b.cnt), _p(b.rowptr), _p(b.pair_c), _p(b.pair_j), _p(ws["x"][l]),
               _p(ws["proj"][l]), _p(ws["g"][l]), eng.w(fg, D * D), eng.w(f"{la}/key/kernel"), eng.w(f"{la}/key/bias"),
               eng.w(f"{la}/layer_norm_g/gamma"), eng.w(f"{la}/layer_norm_g/beta"), eng.w(f"{la}/layer_norm/gamma"),
               eng.w(f"{la}/layer_norm/beta"), _p(ws["g"][l + 1]), _p(ws["ctxpre"][l]), _p(ws["h"][l]), 0)
    scratch_k = torch.empty_like(ws["kk"][l]); scratch_p = torch.empty_like(ws["pre"][l])
    def fwd(which):
        (ntiles, _a0, _a1, _c, _r, pc, pj, x, proj, g_in, W2, Wk, bk, gg, bg, gam, bet, g_out, ctxpre, out, attn) = la_args
        if pipe:
            check(lib.scann_la_forward_pipe(eng.la_grid, b.rows, which, ntiles, pc, pj, x, proj, g_in, W2, Wk, bk, gg, bg, gam, bet,
                                            g_out, ctxpre, out, attn, _p(scratch_p), _p(scratch_k), 0, 0, _p(eng.status), st))
        else:
            check(lib.scann_la_forward_tc(eng.la_grid, b.stride, b.mma_rows, *la_args, _p(scratch_p), _p(scratch_k), 0, 0, st))
    def timeit(fn, n=40):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    res = {}
    for pdl in (0, 1):
        lib.scann_set_pdl(pdl)
        if pipe:
            res[f"geom fwd pdl{pdl}"] = timeit(lambda: fwd(1))
            res[f"attn fwd pdl{pdl}"] = timeit(lambda: fwd(2))
        res[f"geom+attn fwd pdl{pdl}"] = timeit(lambda: fwd(3))
    lib.scann_set_pdl(0)
    eng.check_status()
    print(f"{shape} B={B} pipe={pipe} stride={b.stride} tiles={int(b.ntiles.item())}: " + ", ".join(f"{k} {v:.1f} us" for k, v in res.items()))

run(15, 0)
run(0, 32)
run(0, 64)
