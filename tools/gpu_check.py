"""Stage-by-stage parity dump against the CPU oracle (development tool, GPU only)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200.config import load_yaml, fill_cli_defaults, model_spec, ModelSpec
from scann_b200.params import ParamLayout
from scann_b200.synth import make_batch, count_valid
from scann_b200.engine import Engine
from oracle import scann_oracle as O

QM9 = dict(n_atoms=10, embedding_dim=48, n_attention=7, local_dim=128, num_head=8, global_dim=128, dense_out=128,
           use_attn_norm=True, use_ga_norm=True, use_ring=False, g_update=True, gaussian_d=4.0, feature="atomic",
           use_drop=False, target="homo")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main(shape="qm9", B=8, L=None, seed=0):
    d = dict(QM9)
    if shape == "mp2018":
        d.update(n_atoms=95, embedding_dim=128, n_attention=9, gaussian_d=6.0)
    if L is not None:
        d["n_attention"] = L
    spec = ModelSpec(**d)
    lay = ParamLayout(spec)
    arena = lay.randomize_arena(2)
    w = lay.to_dict(arena)
    inp, tgt = make_batch(shape, seed, B=B)
    print("shape", shape, "B", B, "valid atoms/pairs", count_valid(inp))
    kw = dict(n_attention=spec.n_attention, g_update=True, gaussian_d=spec.gaussian_d, use_attn_norm=True,
              use_ga_norm=spec.use_ga_norm)
    wt = {k: torch.tensor(v, dtype=torch.float64) for k, v in w.items()}
    with torch.no_grad():
        y_ref, ga_ref, tr = O.forward(wt, O.to_torch_inputs(inp), return_all=True, **kw)
    eng = Engine(spec, arena)
    b = eng.load_batch(inp)
    torch.cuda.synchronize()
    eng.check_status()
    print("ntiles", int(b.ntiles.item()), "cap", b.tile_cap, "P", b.P_host)
    cnt = b.cnt.cpu().numpy()
    assert (cnt == inp["neighbor_mask"].reshape(b.R, -1).sum(1)).all(), "cnt mismatch"
    pc = b.pair_c.cpu().numpy(); slot = b.pair_slot.cpu().numpy(); pj = b.pair_j.cpu().numpy()
    valid = pc >= 0
    assert valid.sum() == b.P_host, (valid.sum(), b.P_host)
    # gather / mask bit-exactness
    nbr_flat = (np.arange(b.B)[:, None, None] * b.M + inp["neighbors"]).reshape(-1)
    assert (pj[valid] == nbr_flat[slot[valid]]).all(), "pair_j mismatch"
    assert (pc[valid] == slot[valid] // b.N).all(), "pair_c mismatch"
    assert inp["neighbor_mask"].reshape(-1)[slot[valid]].all()
    print("plan: gathers and masks bit-exact")
    y, ga = eng.forward(b, training=True)
    torch.cuda.synchronize()
    eng.check_status()
    ws = eng._workspace(b, True)
    R = b.R
    print("x0     ", rel(ws["x"][0].cpu().numpy(), tr["x0"].reshape(R, -1)))
    g0 = tr["g0"].reshape(-1, 128).numpy()
    print("g0     ", rel(ws["g"][0].cpu().numpy()[valid], g0[slot[valid]]))
    am = inp["atom_mask"].reshape(-1)
    for l in range(spec.n_attention):
        gl = tr[f"g{l+1}"].reshape(-1, 128).numpy()
        print(f"L{l} ctxLN", rel(ws["h"][l].cpu().numpy(), tr[f"ctx{l}"].reshape(R, -1)),
              "x", rel(ws["x"][l + 1].cpu().numpy(), tr[f"x{l+1}"].reshape(R, -1)),
              "g", rel(ws["g"][l + 1].cpu().numpy()[valid], gl[slot[valid]]))
    print("y      ", rel(y.cpu().numpy(), y_ref.numpy().ravel()), y.cpu().numpy()[:4], y_ref.numpy().ravel()[:4])
    print("ga     ", rel(ga.cpu().numpy(), ga_ref.numpy().ravel()))
    # fp32 oracle noise floor
    y32, ga32 = O.predict(w, inp, torch.float32, **kw)
    print("oracle fp32 vs fp64: y", rel(y32, y_ref.numpy()), "ga", rel(ga32, ga_ref.numpy()))
    # gradients
    l2n = [e.name for e in lay if e.l2]
    loss_ref, _, _, grads_ref = O.loss_and_grads(w, inp, tgt, l2n, **kw)
    tgt_d = torch.from_numpy(tgt).cuda()
    eng.train_step(b, tgt_d, lr=1e-3, apply=False, want_grads=True)
    torch.cuda.synchronize()
    eng.check_status()
    lv = eng.loss_value(b.B).cpu().numpy()
    print("loss", lv[0], "ref", loss_ref, "rel", abs(lv[0] - loss_ref) / abs(loss_ref))
    g = lay.to_dict(eng.grad_out.cpu().numpy())
    worst = []
    for e in lay:
        r = rel(g[e.name], grads_ref[e.name])
        worst.append((r, e.name, float(np.abs(grads_ref[e.name]).max())))
    worst.sort(reverse=True)
    for r, n, m in worst[:12]:
        print(f"grad {n:45s} rel {r:.3e}  |ref|max {m:.3e}")
    print("max grad rel err", worst[0][0])
    # timing
    for name, fn in (("forward", lambda: eng.forward(b, training=False)),
                     ("train_step", lambda: eng.train_step(b, tgt_d, lr=1e-3))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(name, "ms", e0.elapsed_time(e1) / 10)


if __name__ == "__main__":
    args = sys.argv[1:]
    shape = args[0] if args else "qm9"
    B = int(args[1]) if len(args) > 1 else 8
    L = int(args[2]) if len(args) > 2 else None
    main(shape, B, L)
