"""The hyper-parameters of the reference's shipped yaml files as Python dicts.

Values restate configs/model_qm9.yaml, model_qm9_std.yaml, model_mp2018.yaml, model_smfe.yaml,
model_fullerene.yaml and model_ptgp.yaml of the reference -- every yaml it ships (model section and the
hyper-parameters the training shell reads; dataset paths are not reproduced).
A user's own yaml loads unchanged through ``scann_b200.config.load_yaml``.
"""
from __future__ import annotations

import copy

_COMMON = {"local_dim": 128, "num_head": 8, "global_dim": 128, "dense_out": 128, "scale": 0.5,
           "use_attn_norm": True}

CONFIGS = {
    # configs/model_qm9.yaml
    "qm9": {"model": dict(_COMMON, n_atoms=10, embedding_dim=48, n_attention=7, use_ga_norm=True, use_ring=False,
                          g_update=True, gaussian_d=4.0),
            "hyper": {"batch_size": 128, "scaler": True, "scheduler": "sgdr", "lr": 0.0005, "min_lr": 0.0001}},
    # configs/model_qm9_std.yaml (the JCTC split: eight layers, cosine schedule, no scaler key)
    "qm9_std": {"model": dict(_COMMON, n_atoms=10, embedding_dim=48, n_attention=8, use_ga_norm=True, use_ring=False,
                              g_update=True, gaussian_d=4.0),
                "hyper": {"batch_size": 128, "scheduler": "cosine", "lr": 0.0006, "min_lr": 0.00008}},
    # configs/model_mp2018.yaml
    "mp2018": {"model": dict(_COMMON, n_atoms=95, embedding_dim=128, n_attention=9, use_ga_norm=True, use_ring=False,
                             g_update=True, gaussian_d=6.0),
               "hyper": {"batch_size": 64, "scaler": False, "scheduler": "cosine", "lr": 0.0001, "min_lr": 0.00005}},
    # configs/model_smfe.yaml (SmFe12 crystals)
    "smfe": {"model": dict(_COMMON, n_atoms=70, embedding_dim=48, n_attention=9, use_ga_norm=True, use_ring=False,
                           g_update=True, gaussian_d=6.0),
             "hyper": {"batch_size": 128, "scaler": False, "scheduler": "cosine", "lr": 0.0005, "min_lr": 0.0001}},
    # configs/model_fullerene.yaml
    "fullerene": {"model": dict(_COMMON, n_atoms=10, embedding_dim=48, n_attention=7, use_ga_norm=False,
                                use_ring=False, g_update=True, gaussian_d=4.0),
                  "hyper": {"batch_size": 128, "scaler": True, "scheduler": "cosine", "lr": 0.0001, "min_lr": 0.00001}},
    # configs/model_ptgp.yaml (lacks g_update / gaussian_d as shipped -> KeyError in create_model, as in the reference)
    "ptgp": {"model": dict(_COMMON, n_atoms=80, embedding_dim=48, n_attention=11, use_ga_norm=True, use_ring=True),
             "hyper": {"batch_size": 64, "lr": 0.0005, "min_lr": 0.0001}},
}


def get_config(name: str, target: str = "homo", feature: str = "atomic", use_drop: bool = False) -> dict:
    """Config dict as train.py hands it to SCANN: yaml + the CLI-injected keys (train.py:37-43)."""
    cfg = copy.deepcopy(CONFIGS[name])
    cfg["model"]["feature"] = feature
    cfg["model"]["use_drop"] = use_drop
    cfg["hyper"]["target"] = target
    cfg["hyper"]["use_ref"] = False
    return cfg
