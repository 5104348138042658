"""Synthetic padded batches with the layout ``DataIterator.__getitem__`` produces
(reference: scann/utils/datagenerator.py:69-135; single structure:
scann/utils/general.py:206-246).

No dataset is reachable offline, so every benchmark / parity input is generated here
from ``numpy.random.default_rng(seed)``.  Layout contract (same keys, dtypes and
padding rules as the reference):

* ``atomic``            [B,M]   int32, 0 = padded atom (``atom_mask = atomic != 0``)
* ``atom_mask``         [B,M,1] bool
* ``neighbors``         [B,M,N] int32, padded slots reset to index 0
* ``neighbor_mask``     [B,M,N] bool
* ``neighbor_weight``   [B,M,N] float32, zero padded (raw solid angle when g_update)
* ``neighbor_distance`` [B,M,N] float32, zero padded
* ``ring_aromatic``     [B,M,2] int32 (only when use_ring)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class Shape:
    name: str
    B: int
    M: int
    N: int
    atoms: Tuple[int, int]          # inclusive range of atoms / structure
    z_choices: Optional[Sequence[int]]
    z_range: Optional[Tuple[int, int]]
    nbrs: Tuple[int, int]           # inclusive range of neighbours / atom
    dist: Tuple[float, float]
    weight: Tuple[float, float]
    atoms_choices: Optional[Sequence[int]] = None
    config: str = ""


# Canonical synthetic shapes.
SHAPES: Dict[str, Shape] = {
    "qm9": Shape("qm9", 128, 29, 16, (9, 29), (1, 6, 7, 8, 9), None, (2, 16), (0.9, 4.0), (0.4, 3.0),
                 config="model_qm9.yaml"),
    "mp2018": Shape("mp2018", 64, 64, 24, (2, 64), None, (1, 94), (6, 24), (1.5, 6.0), (0.4, 3.0),
                    config="model_mp2018.yaml"),
    "fullerene": Shape("fullerene", 128, 72, 12, (60, 72), (6,), None, (3, 12), (1.3, 4.0), (0.4, 3.0),
                       atoms_choices=(60, 70, 72), config="model_fullerene.yaml"),
    "ptgp": Shape("ptgp", 64, 256, 16, (200, 256), (6, 78), None, (2, 16), (1.0, 4.0), (0.2, 1.0),
                  config="model_ptgp.yaml"),
}


def cgcnn_table(n_species: int = 100, seed: int = 1234) -> np.ndarray:
    """Stand-in for the reference's ``atomic_features`` JSON (scann/utils/dataset.py): one 92-vector per atomic
    number (0/1 entries like the CGCNN one-hot blocks); row 0 (padding) is all zeros."""
    t = (np.random.default_rng(seed).random((n_species, 92)) < 0.15).astype(np.float32)
    t[0] = 0.0
    return t


def make_batch(shape: Shape | str, seed: int = 0, B: Optional[int] = None, use_ring: bool = False,
               full: bool = False, feature: str = "atomic") -> Tuple[Dict[str, np.ndarray], np.ndarray]:
    """Return ``(inputs, target)`` like ``DataIterator.__getitem__``.

    ``full=True`` fills every atom / neighbour slot (the upper-bound shape BASELINE.md
    section 3 quotes); otherwise counts are ragged as in real data.  Neighbour ids are
    uniform over the structure's own atoms, duplicates and self allowed (periodic images
    map to the same site in the reference, voronoi_neighbor.py:42).
    """
    if isinstance(shape, str):
        shape = SHAPES[shape]
    B = shape.B if B is None else int(B)
    M, N = shape.M, shape.N
    rng = np.random.default_rng(seed)

    atomic = np.zeros((B, M), np.int32)
    neighbors = np.zeros((B, M, N), np.int32)
    nmask = np.zeros((B, M, N), bool)
    weight = np.zeros((B, M, N), np.float32)
    dist = np.zeros((B, M, N), np.float32)

    if full:
        n_at = np.full(B, M)
    elif shape.atoms_choices is not None:
        n_at = rng.choice(np.asarray(shape.atoms_choices), size=B)
    else:
        n_at = rng.integers(shape.atoms[0], shape.atoms[1] + 1, size=B)
    if not full and B > 1:
        n_at[rng.integers(0, B)] = M   # DataIterator pads to the longest structure in the batch

    for b in range(B):
        a = int(n_at[b])
        if shape.z_choices is not None:
            atomic[b, :a] = rng.choice(np.asarray(shape.z_choices, np.int32), size=a)
        else:
            atomic[b, :a] = rng.integers(shape.z_range[0], shape.z_range[1] + 1, size=a)
        if full:
            cnt = np.full(a, N)
        else:
            cnt = rng.integers(shape.nbrs[0], shape.nbrs[1] + 1, size=a)
            cnt = np.minimum(cnt, N)
        valid = np.arange(N)[None, :] < cnt[:, None]
        nmask[b, :a] = valid
        neighbors[b, :a] = np.where(valid, rng.integers(0, a, size=(a, N)), 0)
        weight[b, :a] = np.where(valid, rng.uniform(*shape.weight, size=(a, N)), 0.0)
        dist[b, :a] = np.where(valid, rng.uniform(*shape.dist, size=(a, N)), 0.0)
    if not full and B > 1:
        # at least one atom of the longest structure uses every neighbour slot
        b = int(np.argmax(n_at))
        nmask[b, 0] = True
        neighbors[b, 0] = rng.integers(0, int(n_at[b]), size=N)
        weight[b, 0] = rng.uniform(*shape.weight, size=N)
        dist[b, 0] = rng.uniform(*shape.dist, size=N)

    inputs = {
        "atomic": atomic,
        "atom_mask": (atomic != 0)[..., None],
        "neighbors": neighbors,
        "neighbor_mask": nmask,
        "neighbor_weight": weight,
        "neighbor_distance": dist,
    }
    if use_ring:
        inputs["ring_aromatic"] = (rng.integers(0, 2, size=(B, M, 2)) * (atomic != 0)[..., None]).astype(np.int32)
    if feature == "cgcnn":          # DataIterator replaces the atomic numbers by their feature vectors (:109-110)
        inputs["atomic"] = cgcnn_table()[atomic]
    target = rng.standard_normal(B).astype(np.float32)
    return inputs, target


def count_valid(inputs: Dict[str, np.ndarray]) -> Tuple[int, int]:
    """(valid atoms A, valid atom-neighbour pairs P) -- the units algorithmic bytes are counted in."""
    return int(inputs["atom_mask"].sum()), int(inputs["neighbor_mask"].sum())
