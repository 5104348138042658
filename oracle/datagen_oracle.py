"""CPU restatement of the reference's batch assembly -- TEST INFRASTRUCTURE ONLY (imported by tests/).

Follows ``DataIterator.__getitem__`` (scann/utils/datagenerator.py:69-135) and ``pad_sequence`` /
``pad_nested_sequences`` (scann/utils/general.py:14-50) with plain Python loops: neighbour lists are padded with
1000, the mask is ``!= 1000`` and padded indices are reset to 0 (:82-90); weights / distances are zero padded;
``atom_mask = Z != 0`` (:103-107); weight column 3 (normalised solid angle) unless g_update -> column 2 (:48-50).
Parity unpinned at the TensorFlow boundary like the rest of the oracle (no reference fixtures exist); this part
only uses numpy in the reference, so the restatement is exact by construction.
"""
from __future__ import annotations

import numpy as np


def pad_sequence(sequences, maxlen, dtype, value=0):                      # general.py:14-33 (padding="post")
    sample_shape = np.asarray(sequences[0]).shape[1:]
    x = np.full((len(sequences), maxlen) + sample_shape, value, dtype=dtype)
    for i, s in enumerate(sequences):
        t = np.asarray(s[-maxlen:], dtype=dtype)
        x[i, :len(t)] = t
    return x


def pad_nested_sequences(sequences, max_len_1, max_len_2, dtype, value=0):  # general.py:36-50
    inner = [pad_sequence(sq, max_len_1, dtype, value) for sq in sequences]
    return pad_sequence(inner, max_len_2, dtype, value)


def get_item(data_energy, data_neighbor, indexes, *, g_update: bool, use_ring: bool = False, converter: float = 1.0):
    batch_nei = [data_neighbor[i] for i in indexes]
    batch_atom = [data_energy[i] for i in indexes]
    max_c = max(len(c) for c in batch_nei)                                 # :75
    max_n = max(len(n) for c in batch_nei for n in c)                      # :76
    energy = np.array([float(p[1]) * converter for p in batch_atom], "float32")
    wi = 2 if g_update else 3                                              # :48-50
    pad_local = pad_nested_sequences([[[n[1] for n in lc] for lc in p] for p in batch_nei], max_n, max_c, "int32", 1000)
    mask_local = pad_local != 1000                                         # :89
    pad_local[pad_local == 1000] = 0                                       # :90
    w = pad_nested_sequences([[[n[wi] for n in lc] for lc in p] for p in batch_nei], max_n, max_c, "float32")
    d = pad_nested_sequences([[[n[-1] for n in lc] for lc in p] for p in batch_nei], max_n, max_c, "float32")
    pad_atom = pad_sequence([c[0] for c in batch_atom], max_c, "int32", 0)  # :104-105
    inputs = {"atomic": pad_atom, "atom_mask": np.expand_dims(pad_atom != 0, -1), "neighbors": pad_local,
              "neighbor_mask": mask_local, "neighbor_weight": w, "neighbor_distance": d}
    if use_ring:
        inputs["ring_aromatic"] = pad_sequence([c[2] for c in batch_atom], max_c, "int32", 0)
    return inputs, energy
