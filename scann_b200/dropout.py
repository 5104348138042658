"""Host replica of the device dropout mask (``drop_hash`` / ``drop_mult`` in csrc/common.cuh).

Training-mode ``keras.layers.Dropout`` of the reference: rate 0.1 after ``dense_embed``
(scann/models/scann_model.py:374) and inside every ``ResidualNorm`` (scann/layers/attention.py:25-31).
The kernels derive the keep decision of element ``row * 128 + column`` of dropout site ``site`` from a hash of
(seed, site, index), so nothing is stored between forward and backward; this module rebuilds the same masks
with numpy (tests inject them into the oracle).  Sites: 0 = dense_embed, 1 + l = residual_norm layer l."""
from __future__ import annotations

import numpy as np

SITE_DENSE_EMBED = 0


def site_residual_norm(layer: int) -> int:
    return 1 + layer


def threshold(rate: float) -> int:
    return min(int(rate * 4294967296.0), 0xFFFFFFFF)


def drop_hash(seed: int, site: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = idx.astype(np.uint32) * np.uint32(0x9E3779B1)
        x ^= np.uint32((seed + site * 0x85EBCA6B) & 0xFFFFFFFF)
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return x


def drop_mask(seed: int, site: int, rows: int, rate: float, cols: int = 128) -> np.ndarray:
    """[rows, cols] float32 multipliers: 0 or 1/(1-rate) (inverted dropout, as Keras)."""
    idx = np.arange(rows * cols, dtype=np.uint64).astype(np.uint32)
    keep = drop_hash(seed, site, idx) >= np.uint32(threshold(rate))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(rate))
    return np.where(keep, scale, np.float32(0.0)).astype(np.float32).reshape(rows, cols)
