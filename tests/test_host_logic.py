"""CPU tests of the host logic: parameter layout, configs, synthetic batches, lr schedules."""
import numpy as np
import pytest
import yaml

from scann_b200.config import fill_cli_defaults, load_yaml, model_spec
from scann_b200.configs import CONFIGS, get_config
from scann_b200.params import ParamLayout
from scann_b200.synth import SHAPES, count_valid, make_batch


@pytest.mark.parametrize("name,expected", [("qm9", 890977), ("mp2018", 1145089), ("fullerene", 890977)])
def test_param_counts_match_reference_models(name, expected):
    """Trainable-parameter totals of the reference graphs."""
    lay = ParamLayout(model_spec(get_config(name)))
    assert lay.n_params == expected
    assert all(e.offset % 4 == 0 for e in lay)


def test_l2_kernels_are_the_regularised_ones():
    """5 per LA+ResidualNorm block + after_Lc + GA query/key + bf_property."""
    spec = model_spec(get_config("qm9"))
    lay = ParamLayout(spec)
    l2 = [e.name for e in lay if e.l2]
    assert len(l2) == 5 * spec.n_attention + 4
    assert "dense_embed/kernel" not in l2 and "predict_property/kernel" not in l2
    assert "neighbor_d/kernel" not in l2 and "embed_atom/embeddings" not in l2


def test_keras_weight_order_of_local_attention():
    lay = ParamLayout(model_spec(get_config("qm9")))
    names = [e.name for e in lay if e.name.startswith("local_attention/")]
    assert names == ["local_attention/query/kernel", "local_attention/query/bias", "local_attention/key/kernel",
                     "local_attention/key/bias", "local_attention/filter_geo/kernel",
                     "local_attention/filter_geo/bias", "local_attention/layer_norm/gamma",
                     "local_attention/layer_norm/beta", "local_attention/layer_norm_g/gamma",
                     "local_attention/layer_norm_g/beta"]
    assert lay["local_attention/filter_geo/kernel"].shape == (384, 128)
    assert "local_attention_6/key/kernel" in lay and "local_attention_7/key/kernel" not in lay


def test_yaml_roundtrip_and_missing_keys(tmp_path):
    cfg = get_config("qm9")
    p = tmp_path / "model_qm9.yaml"
    p.write_text(yaml.safe_dump({"model": {k: v for k, v in cfg["model"].items() if k not in ("feature", "use_drop")},
                                 "hyper": {k: v for k, v in cfg["hyper"].items() if k not in ("target", "use_ref")}}))
    raw = load_yaml(str(p))
    with pytest.raises(KeyError):            # feature/target come from the CLI in the reference (train.py:37-43)
        model_spec(raw)
    spec = model_spec(fill_cli_defaults(raw))
    assert spec.n_attention == 7 and spec.g_update and spec.gaussian_d == 4.0
    with pytest.raises(KeyError):            # model_ptgp.yaml lacks g_update, as shipped
        model_spec(get_config("ptgp"))


def test_arena_roundtrip():
    lay = ParamLayout(model_spec(get_config("mp2018")))
    arena = lay.randomize_arena(5)
    d = lay.to_dict(arena)
    assert np.array_equal(lay.from_dict(d), arena)
    assert abs(d["local_attention_3/layer_norm/gamma"].mean() - 1.0) < 0.2


@pytest.mark.parametrize("shape", list(SHAPES))
def test_synthetic_batch_layout(shape):
    inputs, target = make_batch(shape, 0, B=4)
    s = SHAPES[shape]
    assert inputs["atomic"].shape == (4, s.M) and inputs["atomic"].dtype == np.int32
    assert inputs["atom_mask"].shape == (4, s.M, 1) and inputs["atom_mask"].dtype == bool
    assert inputs["neighbors"].shape == (4, s.M, s.N)
    nm = inputs["neighbor_mask"]
    # padded slots are reset to index 0 and carry zero weight/distance (datagenerator.py:82-101)
    assert (inputs["neighbors"][~nm] == 0).all()
    assert (inputs["neighbor_weight"][~nm] == 0).all() and (inputs["neighbor_distance"][~nm] == 0).all()
    # neighbours only point at real atoms of the same structure
    n_at = inputs["atom_mask"].sum((1, 2))
    assert (inputs["neighbors"] < n_at[:, None, None]).all()
    assert not nm[~inputs["atom_mask"][..., 0]].any()
    A, P = count_valid(inputs)
    assert A == inputs["atom_mask"].sum() and P == nm.sum() and target.shape == (4,)


def test_cosine_decay_matches_keras_formula():
    from scann_b200.model import CosineDecay
    sch = CosineDecay(1e-4, 100, alpha=0.5)
    assert sch(0) == pytest.approx(1e-4)
    assert sch(100) == pytest.approx(0.5e-4)
    assert sch(1000) == pytest.approx(0.5e-4)
    assert sch(50) == pytest.approx(1e-4 * (0.5 * 0.5 + 0.5))


def test_model_config_round_trip_for_every_shipped_config():
    """model.save writes a model_config from which load_model(path) rebuilds the graph without a yaml
    (scann_model.py:85-96); the JSON uses the reference's layer classes, names and get_config keys."""
    import json
    from scann_b200.config import model_spec
    from scann_b200.configs import get_config
    from scann_b200.model import keras_model_config, spec_from_model_config
    for name in ("qm9", "mp2018", "fullerene", "ptgp"):
        for feature in ("atomic", "cgcnn"):
            cfg = get_config(name, feature=feature, target="e_b" if name == "mp2018" else "homo")
            if name == "ptgp":
                cfg["model"].update(g_update=False, gaussian_d=4.0)
            spec = model_spec(cfg)
            js = keras_model_config(spec)
            back = spec_from_model_config(js)
            if feature == "cgcnn":
                spec = spec.__class__(**{**spec.__dict__, "n_atoms": 0})      # the Dense embedding has no vocabulary
            assert back == spec, (name, feature, back, spec)
            layers = json.loads(js)["config"]["layers"]
            names = [l["name"] for l in layers]
            assert names.count("global_attention") == 1 and "after_Lc" in names
            assert sum(l["class_name"] == "LocalAttention" for l in layers) == spec.n_attention
            la = next(l for l in layers if l["class_name"] == "LocalAttention")["config"]
            assert set(la) >= {"dim", "num_head", "v_proj", "scale", "kq_proj", "dropout", "g_update"}   # attention.py:218-231


def test_reference_layer_package_exports():
    import scann.layers as L
    for n in ("GlobalAttention", "LocalAttention", "ResidualNorm", "GaussianExpansion", "SGDRC", "root_mean_squared_error",
              "r2_square", "gather_shape", "mrelu"):
        assert n in L.__all__ and n in L._CUSTOM_OBJECTS


def test_balanced_shards_equal_counts_and_nearly_equal_pair_counts():
    from scann_b200.dist import balanced_shards
    from scann_b200.synth import make_batch
    for shape, B in (("qm9", 128), ("mp2018", 64)):
        inp, _ = make_batch(shape, seed=0, B=B * 8)
        cost = inp["neighbor_mask"].reshape(B * 8, -1).sum(1)
        shards = balanced_shards(cost, 8)
        assert sorted(np.concatenate(shards).tolist()) == list(range(B * 8))        # a partition
        assert all(len(s) == B for s in shards)
        loads = np.array([cost[s].sum() for s in shards])
        assert loads.max() - loads.min() <= 0.005 * loads.mean()                     # within 0.5 %
        naive = np.array([cost[r * B:(r + 1) * B].sum() for r in range(8)])
        assert naive.max() - naive.min() > 4 * (loads.max() - loads.min())
    with pytest.raises(ValueError):
        balanced_shards(np.ones(10), 4)


def test_dlpack_import_accepts_capsules_and_exporters():
    """scann_b200.dlpack: the two DLPack producer forms (capsule = what tf.experimental.dlpack.to_dlpack returns,
    ``__dlpack__`` exporter) come back as zero-copy torch views; numpy arrays and tensors pass through untouched."""
    import torch
    from scann_b200.dlpack import export_capsule, import_tensor, is_capsule

    class Foreign:                                   # an exporter that is neither numpy nor torch
        def __init__(self, t): self.t = t
        def __dlpack__(self, **kw): return self.t.__dlpack__(**kw)
        def __dlpack_device__(self): return self.t.__dlpack_device__()

    src = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    cap = export_capsule(src)
    assert is_capsule(cap)
    for got in (import_tensor(cap), import_tensor(Foreign(src))):
        assert isinstance(got, torch.Tensor) and got.shape == src.shape and got.data_ptr() == src.data_ptr()
    a = np.zeros(3, np.float32)
    assert import_tensor(a) is a and import_tensor(src) is src and import_tensor(None) is None
