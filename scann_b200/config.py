"""Config handling for the SCANN hot path.

The reference reads a yaml dict with two sections, ``model`` and ``hyper``
(reference: configs/*.yaml; consumed at scann/models/scann_model.py:329-447 and
train.py:37-43).  The same dicts load here unchanged.  Keys the reference's CLI
injects (``feature``, ``use_drop``, ``target``, ``use_ref``; train.py:37-43) are
not in the yaml files; the reference has no default for them and raises
``KeyError``.  We keep that behaviour for keys that change arithmetic
(``g_update``, ``gaussian_d``) unless the caller opts into ``fill_cli_defaults``,
which applies the CLI's own argparse defaults (train.py:63-100).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import yaml

D_MODEL = 128  # every shipped config uses local_dim = global_dim = dense_out = 128
N_HEAD = 8
N_RBF = 20


@dataclass(frozen=True)
class ModelSpec:
    """The arithmetic-relevant subset of ``config['model']`` / ``config['hyper']``."""

    n_atoms: int
    embedding_dim: int
    n_attention: int
    local_dim: int
    num_head: int
    global_dim: int
    dense_out: int
    use_attn_norm: bool
    use_ga_norm: bool
    use_ring: bool
    g_update: bool
    gaussian_d: float
    feature: str
    use_drop: bool
    target: str

    @property
    def mrelu_head(self) -> bool:
        # scann_model.py:446 -- ReLU-with-identity-gradient only for target "e_b"
        return self.target == "e_b"


def load_yaml(path: str) -> dict:
    with open(path, "r") as f:
        return yaml.safe_load(f)


def fill_cli_defaults(config: dict) -> dict:
    """Apply what train.py's argparse would inject (train.py:37-43, 63-100)."""
    cfg = copy.deepcopy(config)
    cfg.setdefault("model", {})
    cfg.setdefault("hyper", {})
    cfg["model"].setdefault("feature", "atomic")
    cfg["model"].setdefault("use_drop", False)
    cfg["hyper"].setdefault("target", "homo")
    cfg["hyper"].setdefault("use_ref", False)
    return cfg


def model_spec(config: dict) -> ModelSpec:
    """Read exactly the keys ``create_model`` reads (scann_model.py:330-447).

    Missing keys raise ``KeyError`` as in the reference (e.g. model_ptgp.yaml has
    no ``g_update``/``gaussian_d`` and fails in ``create_model`` as shipped).
    """
    cfm = config["model"]
    spec = ModelSpec(
        n_atoms=int(cfm["n_atoms"]),
        embedding_dim=int(cfm["embedding_dim"]),
        n_attention=int(cfm["n_attention"]),
        local_dim=int(cfm["local_dim"]),
        num_head=int(cfm["num_head"]),
        global_dim=int(cfm["global_dim"]),
        dense_out=int(cfm["dense_out"]),
        use_attn_norm=bool(cfm["use_attn_norm"]),
        use_ga_norm=bool(cfm["use_ga_norm"]),
        use_ring=bool(cfm["use_ring"]),
        g_update=bool(cfm["g_update"]),
        gaussian_d=float(cfm["gaussian_d"]),
        feature=str(cfm["feature"]),
        use_drop=bool(cfm["use_drop"]),
        target=str(config["hyper"]["target"]),
    )
    if spec.feature not in ("atomic", "cgcnn"):
        raise ValueError(f"feature must be 'atomic' or 'cgcnn', got {spec.feature!r}")
    return spec


def check_kernel_support(spec: ModelSpec) -> None:
    """The sm_100a kernels are specialised for the shapes every shipped config uses."""
    if spec.local_dim != D_MODEL or spec.global_dim != D_MODEL or spec.dense_out != D_MODEL:
        raise NotImplementedError("kernels are specialised for local_dim = global_dim = dense_out = 128")
    if spec.num_head != N_HEAD:
        raise NotImplementedError("kernels are specialised for num_head = 8")
    if spec.embedding_dim > 128:
        raise NotImplementedError("embedding_dim must be <= 128")
