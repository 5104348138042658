"""CPU tests of the oracle: golden vectors and structural invariants of the reference graph
(padding invariance, batch-split invariance, permutation equivariance of ga_score,
sum(ga_score) = 1, fully masked rows contribute exactly 0, algebraic identities the kernels use)."""
import os

import numpy as np
import pytest
import torch

from oracle import scann_oracle as O
from scann_b200.config import model_spec
from scann_b200.configs import get_config
from scann_b200.params import ParamLayout
from scann_b200.synth import make_batch
from tests.golden.make_golden import CASES, build_case, oracle_kwargs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def small_model(L=2, cfg_name="qm9", seed=2):
    cfg = get_config(cfg_name)
    cfg["model"]["n_attention"] = L
    spec = model_spec(cfg)
    lay = ParamLayout(spec)
    return spec, lay, lay.to_dict(lay.randomize_arena(seed))


@pytest.mark.parametrize("name", list(CASES))
def test_golden_forward_and_grads(name):
    cfg, spec, lay, arena, inputs, target = build_case(name)
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    w = lay.to_dict(arena)
    l2n = [e.name for e in lay if e.l2]
    loss, y, ga, grads = O.loss_and_grads(w, inputs, target, l2n, dtype=torch.float64, **oracle_kwargs(spec))
    np.testing.assert_allclose(y, z["y"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ga, z["ga"], rtol=1e-10, atol=1e-14)
    assert abs(loss - float(z["loss"])) < 1e-10
    garena = lay.from_dict({k: v.astype(np.float32) for k, v in grads.items()})
    np.testing.assert_allclose(garena[z["grad_idx"]], z["grad_sample"], rtol=1e-6, atol=1e-9)


def test_fp32_reference_noise_floor():
    """The reference's own arithmetic is fp32: record how far an fp32 run sits from fp64 truth."""
    cfg, spec, lay, arena, inputs, target = build_case("qm9_b4")
    w = lay.to_dict(arena)
    y64, ga64 = O.predict(w, inputs, torch.float64, **oracle_kwargs(spec))
    y32, ga32 = O.predict(w, inputs, torch.float32, **oracle_kwargs(spec))
    assert np.abs(y32 - y64).max() <= 2e-5 * np.abs(y64).max()
    assert np.abs(ga32 - ga64).max() <= 1e-5 * np.abs(ga64).max()


def test_ga_score_sums_to_one_and_masks():
    spec, lay, w = small_model()
    inputs, _ = make_batch("qm9", 3, B=5)
    y, ga = O.predict(w, inputs, **oracle_kwargs(spec))
    np.testing.assert_allclose(ga.sum(1).ravel(), 1.0, rtol=1e-12)
    assert (ga[~inputs["atom_mask"]] == 0).all()          # exp(-1e9 - max) underflows to exactly 0


def test_padding_invariance():
    """Extra padded atoms / neighbour slots never change a valid output."""
    spec, lay, w = small_model()
    inputs, _ = make_batch("qm9", 4, B=3)
    y0, ga0 = O.predict(w, inputs, **oracle_kwargs(spec))
    B, M, N = inputs["neighbors"].shape
    pad = {}
    for k, v in inputs.items():
        if v.ndim == 3 and v.shape[2] == N:
            pad[k] = np.pad(v, ((0, 0), (0, 3), (0, 2)))
        elif v.ndim == 3:
            pad[k] = np.pad(v, ((0, 0), (0, 3), (0, 0)))
        else:
            pad[k] = np.pad(v, ((0, 0), (0, 3)))
    y1, ga1 = O.predict(w, pad, **oracle_kwargs(spec))
    np.testing.assert_allclose(y1, y0, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ga1[:, :M], ga0, rtol=1e-12, atol=1e-16)
    assert (ga1[:, M:] == 0).all()


def test_batch_split_invariance():
    """Structures are independent: sharding a batch (data parallelism) never changes results."""
    spec, lay, w = small_model()
    inputs, _ = make_batch("qm9", 5, B=6)
    y, ga = O.predict(w, inputs, **oracle_kwargs(spec))
    for lo, hi in ((0, 2), (2, 6)):
        ys, gas = O.predict(w, {k: v[lo:hi] for k, v in inputs.items()}, **oracle_kwargs(spec))
        np.testing.assert_allclose(ys, y[lo:hi], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(gas, ga[lo:hi], rtol=1e-12, atol=1e-16)


def test_atom_permutation_equivariance():
    spec, lay, w = small_model()
    inputs, _ = make_batch("qm9", 6, B=2)
    B, M, N = inputs["neighbors"].shape
    y0, ga0 = O.predict(w, inputs, **oracle_kwargs(spec))
    rng = np.random.default_rng(0)
    perm = rng.permutation(M)                 # new position p holds old atom perm[p]
    inv = np.argsort(perm)
    out = {}
    for k, v in inputs.items():
        out[k] = v[:, perm].copy()
    out["neighbors"] = np.where(out["neighbor_mask"], inv[out["neighbors"]], 0).astype(np.int32)
    y1, ga1 = O.predict(w, out, **oracle_kwargs(spec))
    np.testing.assert_allclose(y1, y0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ga1, ga0[:, perm], rtol=1e-9, atol=1e-14)


def test_filter_geo_split_identity():
    """[x_c | g | nbr] @ Wf == x_c@W1 + g@W2 + nbr@W3 (the per-atom / per-pair split the kernels use)."""
    rng = np.random.default_rng(0)
    x, g, nb = (torch.tensor(rng.standard_normal((7, 128))) for _ in range(3))
    Wf = torch.tensor(rng.standard_normal((384, 128)))
    full = torch.cat([x, g, nb], -1) @ Wf
    split = x @ Wf[:128] + g @ Wf[128:256] + nb @ Wf[256:]
    assert float((full - split).abs().max()) < 1e-12


def test_global_attention_linear_identity():
    """s_i = m_i k_i.(Q - m_i q_i) equals the reference's masked [M,M] energy row sums (attention.py:279-292)."""
    rng = np.random.default_rng(1)
    M = 9
    q, k = torch.tensor(rng.standard_normal((M, 128))), torch.tensor(rng.standard_normal((M, 128)))
    m = torch.tensor((rng.random(M) > 0.3).astype(np.float64)).unsqueeze(-1)
    energy = (m * k) @ (m * q).T
    energy = energy * (1 - torch.eye(M, dtype=torch.float64))
    ref = m.squeeze(-1) * energy.sum(-1)
    Q = (m * q).sum(0)
    mine = m.squeeze(-1) * ((k * (Q - m * q)).sum(-1))
    assert float((ref - mine).abs().max()) < 1e-12


def test_single_atom_structure_is_nan_with_ga_norm():
    """tf.linalg.normalize has no epsilon: a 1-atom structure gives 0/0 = NaN (attention.py:297)."""
    spec, lay, w = small_model()
    inputs, _ = make_batch("qm9", 7, B=2)
    inputs["atomic"][0, 1:] = 0
    inputs["atom_mask"] = (inputs["atomic"] != 0)[..., None]
    inputs["neighbor_mask"][0, 1:] = False
    inputs["neighbors"][0] = 0
    y, ga = O.predict(w, inputs, **oracle_kwargs(spec))
    assert np.isnan(y[0]).all() and np.isnan(ga[0]).all()
    assert np.isfinite(y[1]).all()


def test_masked_pairs_have_zero_gradient_effect():
    """Changing geometry inputs of masked slots changes neither outputs nor gradients."""
    spec, lay, w = small_model(L=1)
    inputs, target = make_batch("qm9", 8, B=2)
    l2n = [e.name for e in lay if e.l2]
    l0, y0, _, g0 = O.loss_and_grads(w, inputs, target, l2n, **oracle_kwargs(spec))
    alt = {k: v.copy() for k, v in inputs.items()}
    nm = ~alt["neighbor_mask"]
    alt["neighbor_distance"][nm] = 2.5
    alt["neighbor_weight"][nm] = 1.5
    l1, y1, _, g1 = O.loss_and_grads(w, alt, target, l2n, **oracle_kwargs(spec))
    assert abs(l0 - l1) < 1e-13
    for k in g0:
        np.testing.assert_allclose(g1[k], g0[k], rtol=1e-9, atol=1e-13)


def test_adam_legacy_decay():
    p, m, v = np.ones(3), np.zeros(3), np.zeros(3)
    g = np.array([0.1, -0.2, 0.3])
    p1, m1, v1 = O.adam_legacy_step(p, g, m, v, 1, 1e-3)
    # first step of Adam moves every coordinate by ~lr in the direction of -sign(g)
    np.testing.assert_allclose(p1, 1 - 1e-3 * np.sign(g), rtol=1e-5)
    p2, _, _ = O.adam_legacy_step(p1, g, m1, v1, 2, 1e-3)
    assert np.all(np.abs(p2 - p1) < 1e-3 * (1 + 1e-9))       # lr_t = lr / (1 + 1e-5 * 1) < lr
