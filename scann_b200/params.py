"""Parameter layout of the SCANN graph, in the reference's Keras layer / weight order.

Order and shapes follow the attribute-creation order of the reference layers:
``create_model`` (scann/models/scann_model.py:361-447), ``LocalAttention.__init__``
(scann/layers/attention.py:95-113), ``ResidualNorm.__init__`` (:25-35),
``GlobalAttention.__init__`` (:260-262).  Dense kernels are ``[in, out]`` and are
used as ``x @ W + b``.

All parameters live in ONE flat fp32 arena (and one gradient arena of the same
shape): a single Adam kernel and a single gradient all-reduce cover the model.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import numpy as np

from .config import ModelSpec, N_RBF


@dataclass(frozen=True)
class ParamEntry:
    name: str                 # "<layer>/<sublayer>/<weight>"
    shape: Tuple[int, ...]
    offset: int               # element offset in the flat arena
    init: str                 # "glorot" | "zeros" | "ones" | "embed"
    l2: bool                  # carries kernel_regularizer=l2(1e-4) in the reference

    @property
    def size(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def layer_name(base: str, i: int) -> str:
    """Keras auto-naming: first instance has no suffix, then _1, _2, ..."""
    return base if i == 0 else f"{base}_{i}"


class ParamLayout:
    def __init__(self, spec: ModelSpec):
        self.spec = spec
        self.entries: List[ParamEntry] = []
        self._by_name: Dict[str, ParamEntry] = {}
        off = 0

        def add(name, shape, init, l2=False):
            nonlocal off
            e = ParamEntry(name, tuple(int(s) for s in shape), off, init, l2)
            self.entries.append(e)
            self._by_name[name] = e
            # keep every tensor 16-byte aligned for float4 loads
            off += (e.size + 3) // 4 * 4

        D = spec.local_dim
        E = spec.embedding_dim
        if spec.feature == "atomic":
            add("embed_atom/embeddings", (spec.n_atoms, E), "embed")          # scann_model.py:362
        else:
            add("embed_atom/kernel", (92, E), "glorot")                        # scann_model.py:365
            add("embed_atom/bias", (E,), "zeros")
        e_in = E
        if spec.use_ring:
            add("extra_embed/kernel", (2, 10), "glorot")                       # scann_model.py:368
            add("extra_embed/bias", (10,), "zeros")
            e_in = E + 10
        add("dense_embed/kernel", (e_in, D), "glorot")                         # scann_model.py:373
        add("dense_embed/bias", (D,), "zeros")
        if spec.g_update:
            add("neighbor_d/kernel", (N_RBF, D), "glorot")                     # scann_model.py:381
            add("neighbor_d/bias", (D,), "zeros")
            add("neighbor_w/kernel", (N_RBF, D), "glorot")                     # scann_model.py:386
            add("neighbor_w/bias", (D,), "zeros")
        for l in range(spec.n_attention):
            la = layer_name("local_attention", l)
            add(f"{la}/query/kernel", (D, D), "glorot", l2=True)               # attention.py:95
            add(f"{la}/query/bias", (D,), "zeros")
            add(f"{la}/key/kernel", (D, D), "glorot", l2=True)                 # attention.py:97
            add(f"{la}/key/bias", (D,), "zeros")
            k_in = 3 * D if spec.g_update else N_RBF
            add(f"{la}/filter_geo/kernel", (k_in, D), "glorot", l2=True)       # attention.py:104-109
            add(f"{la}/filter_geo/bias", (D,), "zeros")
            add(f"{la}/layer_norm/gamma", (D,), "ones")                        # attention.py:111
            add(f"{la}/layer_norm/beta", (D,), "zeros")
            if spec.g_update:
                add(f"{la}/layer_norm_g/gamma", (D,), "ones")                  # attention.py:113
                add(f"{la}/layer_norm_g/beta", (D,), "zeros")
            if spec.use_attn_norm:
                rn = layer_name("residual_norm", l)
                add(f"{rn}/dense/kernel", (D, D), "glorot", l2=True)           # attention.py:27
                add(f"{rn}/dense/bias", (D,), "zeros")
                add(f"{rn}/dense_1/kernel", (D, D), "glorot", l2=True)         # attention.py:28
                add(f"{rn}/dense_1/bias", (D,), "zeros")
                add(f"{rn}/layer_norm/gamma", (D,), "ones")                    # attention.py:35
                add(f"{rn}/layer_norm/beta", (D,), "zeros")
        G = spec.global_dim
        add("after_Lc/kernel", (D, G), "glorot", l2=True)                      # scann_model.py:424-429
        add("after_Lc/bias", (G,), "zeros")
        add("global_attention/query/kernel", (G, G), "glorot", l2=True)        # attention.py:260
        add("global_attention/query/bias", (G,), "zeros")
        add("global_attention/key/kernel", (G, G), "glorot", l2=True)          # attention.py:262
        add("global_attention/key/bias", (G,), "zeros")
        add("bf_property/kernel", (G, spec.dense_out), "glorot", l2=True)      # scann_model.py:437-442
        add("bf_property/bias", (spec.dense_out,), "zeros")
        add("predict_property/kernel", (spec.dense_out, 1), "glorot")          # scann_model.py:445-447
        add("predict_property/bias", (1,), "zeros")
        self.total = off

    # ------------------------------------------------------------------ access
    def __getitem__(self, name: str) -> ParamEntry:
        return self._by_name[name]

    def __contains__(self, name: str) -> bool:
        return name in self._by_name

    def __iter__(self) -> Iterator[ParamEntry]:
        return iter(self.entries)

    @property
    def n_params(self) -> int:
        """Number of trainable scalars (what Keras' model.summary() prints)."""
        return sum(e.size for e in self.entries)

    def l2_mask(self) -> np.ndarray:
        """1.0 where the element belongs to an l2-regularised kernel, else 0.0."""
        m = np.zeros(self.total, np.float32)
        for e in self.entries:
            if e.l2:
                m[e.offset:e.offset + e.size] = 1.0
        return m

    # ------------------------------------------------------------------ init
    def init_arena(self, seed: int = 1) -> np.ndarray:
        """Keras default initialisers: Glorot-uniform kernels, zero biases,
        LayerNorm gamma=1/beta=0, Embedding U(-0.05, 0.05)."""
        rng = np.random.default_rng(seed)
        arena = np.zeros(self.total, np.float32)
        for e in self.entries:
            view = arena[e.offset:e.offset + e.size]
            if e.init == "glorot":
                fan_in, fan_out = e.shape[0], e.shape[1]
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                view[:] = rng.uniform(-lim, lim, e.size).astype(np.float32)
            elif e.init == "embed":
                view[:] = rng.uniform(-0.05, 0.05, e.size).astype(np.float32)
            elif e.init == "ones":
                view[:] = 1.0
            elif e.init == "zeros":
                view[:] = 0.0
            else:
                raise ValueError(e.init)
        return arena

    def randomize_arena(self, seed: int = 2, scale: float = 0.3) -> np.ndarray:
        """Non-trivial biases / LayerNorm affine terms so parity tests exercise every
        parameter (a freshly initialised model has zero biases and unit gammas)."""
        rng = np.random.default_rng(seed)
        arena = self.init_arena(seed)
        for e in self.entries:
            view = arena[e.offset:e.offset + e.size]
            if e.init == "zeros":
                view[:] = (scale * rng.standard_normal(e.size)).astype(np.float32)
            elif e.init == "ones":
                view[:] = (1.0 + scale * rng.standard_normal(e.size)).astype(np.float32)
        return arena

    def to_dict(self, arena: np.ndarray) -> Dict[str, np.ndarray]:
        return {e.name: arena[e.offset:e.offset + e.size].reshape(e.shape) for e in self.entries}

    def from_dict(self, weights: Dict[str, np.ndarray]) -> np.ndarray:
        arena = np.zeros(self.total, np.float32)
        for e in self.entries:
            w = np.asarray(weights[e.name], np.float32)
            if tuple(w.shape) != e.shape:
                raise ValueError(f"{e.name}: expected {e.shape}, got {tuple(w.shape)}")
            arena[e.offset:e.offset + e.size] = w.reshape(-1)
        return arena
