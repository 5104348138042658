"""DLPack hand-over on the host side of the boundary (BASELINE.json north_star: tensors are exchanged with the
TF / Keras graph via DLPack; INTEGRATION.md section 2 shows the TensorFlow leg).

Two producer forms exist and both are accepted wherever the layer mirrors / the engine take an array:

* a ``PyCapsule`` named ``"dltensor"`` -- what ``tf.experimental.dlpack.to_dlpack(t)`` and
  ``torch.utils.dlpack.to_dlpack(t)`` return (a capsule can be consumed ONCE);
* any object with ``__dlpack__`` / ``__dlpack_device__`` (the array-API protocol: torch, cupy, jax, numpy >= 1.22).

The import is zero-copy: the returned torch tensor aliases the producer's memory, on the producer's device.  The C ABI
itself takes raw device pointers (include/scann_b200.h), so nothing DLPack-specific crosses it.
"""
from __future__ import annotations

import numpy as np
import torch


def is_capsule(x) -> bool:
    return type(x).__name__ == "PyCapsule"


def import_tensor(x):
    """``x`` as a torch tensor when it is a DLPack capsule or a foreign ``__dlpack__`` exporter; numpy arrays, torch
    tensors and everything else are returned unchanged (host arrays take the pinned staging path)."""
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x
    if is_capsule(x) or hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    return x


def export_capsule(t: torch.Tensor):
    """The capsule a consumer (``tf.experimental.dlpack.from_dlpack``) takes; the tensor must stay alive until the
    consumer has imported it, and the producing stream must be synchronised with the consumer's."""
    return torch.utils.dlpack.to_dlpack(t.contiguous())
