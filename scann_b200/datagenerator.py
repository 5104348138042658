"""``DataIterator`` of the reference (scann/utils/datagenerator.py:12-135) on flat CSR arrays.

The reference rebuilds every padded batch from nested Python lists (three list comprehensions over all
atom-neighbour pairs per ``__getitem__``) -- at 10^5 structures/s on the GPU that host work is the bottleneck.
Here the ragged data are converted ONCE into CSR arrays (structure -> atoms -> neighbours); a batch is either

* ``__getitem__(i)``: the same padded dict ``(inputs, target)`` the reference returns, packed with vectorised
  numpy (bit-identical to the reference's padding, tests/test_datagen.py), or
* ``csr_item(i)``: the CSR slice of the batch, which ``Engine.load_batch_csr`` ships in one host->device copy
  (only valid atoms / pairs travel) and ``scann_pack_batch`` (csrc/plan.cu) expands into the padded device
  buffers.

Data format (reference): ``data_neighbor[s][a]`` = list of neighbour tuples ``n`` with ``n[1]`` = neighbour atom
index, ``n[2]`` = solid angle, ``n[3]`` = normalised solid angle, ``n[-1]`` = distance; ``data_energy[s]`` =
``(atomic numbers, target[, ring/aromatic flags per atom])``.
"""
from __future__ import annotations

from math import ceil
from typing import Dict, Tuple

import numpy as np


ATOMIC_FEATURES = None      # [Z + 1, 92] float32 table for feature='cgcnn' (set_atomic_features / load_atomic_features)


def set_atomic_features(table) -> np.ndarray:
    """Install the per-element feature vectors ``feature='cgcnn'`` batches are built from -- the reference keeps them
    as ``atomic_features`` (scann/utils/dataset/atomic_data.py: CGCNN's ``atom_init.json``, one 92-vector per atomic
    number, keys are strings, key "0" is the padding atom).  ``table``: such a dict, or an array ``[Z + 1, 92]``.
    The table itself is data of the CGCNN project and is not shipped with this package."""
    global ATOMIC_FEATURES
    if isinstance(table, dict):
        zmax = max(int(k) for k in table)
        width = len(next(iter(table.values())))
        arr = np.zeros((zmax + 1, width), np.float32)
        for k, v in table.items():
            arr[int(k)] = np.asarray(v, np.float32)
    else:
        arr = np.ascontiguousarray(table, np.float32)
    if arr.ndim != 2:
        raise ValueError("atomic feature table must be [Z + 1, features]")
    ATOMIC_FEATURES = arr
    return arr


def load_atomic_features(path: str) -> np.ndarray:
    """``set_atomic_features`` from a JSON file in CGCNN's ``atom_init.json`` format ({"1": [...], "2": [...], ...})."""
    import json
    with open(path) as f:
        return set_atomic_features(json.load(f))


class DataIterator:
    def __init__(self, data_energy, data_neighbor, batch_size=32, converter=False, use_ring=False, shuffle=False,
                 feature="atomic", g_update=False, atomic_features=None):
        if feature not in ("atomic", "cgcnn"):
            raise ValueError(f"feature must be 'atomic' or 'cgcnn', got {feature!r}")
        if atomic_features is not None:
            set_atomic_features(atomic_features)
        if feature == "cgcnn" and ATOMIC_FEATURES is None:
            raise NotImplementedError("feature='cgcnn' needs the per-element feature table: pass atomic_features= or call "
                                      "scann_b200.datagenerator.load_atomic_features(<atom_init.json>)")
        self.batch_size, self.shuffle, self.use_ring, self.feature = batch_size, shuffle, use_ring, feature
        self.weight_index = 2 if g_update else 3                       # datagenerator.py:48-50
        self.converter = 1000 if converter else 1.0                    # :54-57
        S = len(data_energy)
        n_atoms = np.fromiter((len(c) for c in data_neighbor), np.int64, S)
        self.atom_off = np.zeros(S + 1, np.int64)
        np.cumsum(n_atoms, out=self.atom_off[1:])
        A = int(self.atom_off[-1])
        n_nbr = np.fromiter((len(lc) for c in data_neighbor for lc in c), np.int64, A)
        self.nbr_off = np.zeros(A + 1, np.int64)
        np.cumsum(n_nbr, out=self.nbr_off[1:])
        P = int(self.nbr_off[-1])
        self.nbr_idx = np.fromiter((n[1] for c in data_neighbor for lc in c for n in lc), np.int32, P)
        wi = self.weight_index
        self.nbr_w = np.fromiter((n[wi] for c in data_neighbor for lc in c for n in lc), np.float32, P)
        self.nbr_d = np.fromiter((n[-1] for c in data_neighbor for lc in c for n in lc), np.float32, P)
        self.z = np.fromiter((z for p in data_energy for z in p[0]), np.int32, A)
        self.energy = np.array([float(p[1]) * self.converter for p in data_energy], "float32")
        self.ring = (np.array([r for p in data_energy for r in p[2]], np.int32).reshape(A, 2) if use_ring else None)
        self.on_epoch_end()

    def on_epoch_end(self):
        self.indexes = np.arange(len(self.energy))
        if self.shuffle:
            np.random.shuffle(self.indexes)

    def __len__(self):
        return ceil(len(self.energy) / self.batch_size)

    # ------------------------------------------------------------------ CSR slice of one batch
    def csr_item(self, idx: int) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
        sel = self.indexes[idx * self.batch_size:(idx + 1) * self.batch_size]
        a0, a1 = self.atom_off[sel], self.atom_off[sel + 1]
        n_at = (a1 - a0).astype(np.int64)
        atoms = np.concatenate([np.arange(s, e) for s, e in zip(a0, a1)]) if len(sel) else np.zeros(0, np.int64)
        p0, p1 = self.nbr_off[atoms], self.nbr_off[atoms + 1]
        n_nb = (p1 - p0).astype(np.int64)
        pairs = np.concatenate([np.arange(s, e) for s, e in zip(p0, p1)]) if len(atoms) else np.zeros(0, np.int64)
        sa = np.zeros(len(sel) + 1, np.int32)
        np.cumsum(n_at, out=sa[1:])
        an = np.zeros(len(atoms) + 1, np.int32)
        np.cumsum(n_nb, out=an[1:])
        csr = {"struct_atom_off": sa, "atom_nbr_off": an, "z": self.z[atoms], "nbr_idx": self.nbr_idx[pairs],
               "nbr_w": self.nbr_w[pairs], "nbr_d": self.nbr_d[pairs],
               "M": int(n_at.max()) if len(sel) else 0, "N": int(n_nb.max()) if len(atoms) else 0}
        if self.use_ring:
            csr["ring"] = self.ring[atoms]
        return csr, self.energy[sel]

    # ------------------------------------------------------------------ padded batch, as the reference returns it
    def __getitem__(self, idx: int):
        csr, energy = self.csr_item(idx)
        inputs = pack_padded(csr)
        if self.feature == "cgcnn":        # datagenerator.py:107-110: the mask comes from the atomic numbers, then the
            inputs["atomic"] = ATOMIC_FEATURES[inputs["atomic"]]      # numbers are replaced by their feature vectors
        return inputs, energy


def pack_padded(csr: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """CSR batch -> the padded dict of DataIterator.__getitem__ (vectorised; the device kernel does the same)."""
    sa, an = csr["struct_atom_off"].astype(np.int64), csr["atom_nbr_off"].astype(np.int64)
    B, M, N = len(sa) - 1, csr["M"], csr["N"]
    A = int(sa[-1])
    b_of_atom = np.repeat(np.arange(B), np.diff(sa))
    m_of_atom = np.arange(A) - sa[b_of_atom]
    atomic = np.zeros((B, M), np.int32)
    atomic[b_of_atom, m_of_atom] = csr["z"]
    a_of_pair = np.repeat(np.arange(A), np.diff(an))
    n_of_pair = np.arange(int(an[-1])) - an[a_of_pair]
    bp, mp = b_of_atom[a_of_pair], m_of_atom[a_of_pair]
    nbr = np.zeros((B, M, N), np.int32)
    nmask = np.zeros((B, M, N), bool)
    w = np.zeros((B, M, N), np.float32)
    d = np.zeros((B, M, N), np.float32)
    idx = csr["nbr_idx"]
    nmask[bp, mp, n_of_pair] = idx != 1000          # the reference pads with 1000 and masks "!= 1000" (:82-90)
    nbr[bp, mp, n_of_pair] = np.where(idx == 1000, 0, idx)
    w[bp, mp, n_of_pair] = csr["nbr_w"]
    d[bp, mp, n_of_pair] = csr["nbr_d"]
    out = {"atomic": atomic, "atom_mask": (atomic != 0)[..., None], "neighbors": nbr, "neighbor_mask": nmask,
           "neighbor_weight": w, "neighbor_distance": d}
    if "ring" in csr:
        ring = np.zeros((B, M, 2), np.int32)
        ring[b_of_atom, m_of_atom] = csr["ring"]
        out["ring_aromatic"] = ring
    return out


def pad_sequence(sequences, maxlen=None, dtype="int32", value=0, padding="post"):
    """``pad_sequence`` (scann/utils/general.py:14-32): ``[n, maxlen, ...]`` array filled with ``value``, sequence ``i``
    at the front of row ``i``; a sequence longer than ``maxlen`` keeps its LAST ``maxlen`` items.  ``padding`` is
    accepted and, as in the reference, always behaves as "post"."""
    if maxlen is None:
        maxlen = max(len(seq) for seq in sequences)
    out = np.full((len(sequences), maxlen) + np.asarray(sequences[0]).shape[1:], value, dtype=dtype)
    for row, seq in zip(out, sequences):
        kept = np.asarray(seq[-maxlen:], dtype=dtype)
        row[:len(kept)] = kept
    return out


def pad_nested_sequences(sequences, max_len_1, max_len_2, dtype="int32", value=0):
    """``pad_nested_sequences`` (scann/utils/general.py:35-50): ragged ``[structure][atom][neighbour]`` lists ->
    ``[n, max_len_2 atoms, max_len_1 neighbours]`` filled with ``value``."""
    out = np.full((len(sequences), max_len_2, max_len_1), value, dtype=dtype)
    for block, atoms in zip(out, sequences):
        for row, items in zip(block, atoms[-max_len_2:]):
            kept = np.asarray(items[-max_len_1:], dtype=dtype)
            row[:len(kept)] = kept
    return out


def split_data(len_data, test_percent=0.1, train_size=None, test_size=None):
    """``split_data`` (scann/utils/general.py:79-101): ONE permutation of the global numpy generator (the same
    ``np.random.seed`` gives the reference's split) cut into train | valid | test | extra.  Sizes: the given
    ``train_size`` / ``test_size``, else ``int(len * (1 - 2 * test_percent))`` / ``int(len * test_percent)``; the
    validation set takes what is left."""
    if train_size:
        n_train, n_test = train_size, test_size
    else:
        n_train, n_test = int(len_data * (1 - test_percent * 2)), int(len_data * test_percent)
    n_valid = len_data - n_train - n_test
    order = np.random.permutation(len_data)
    train, valid, test, extra = np.split(order, np.cumsum([n_train, n_valid, n_test]))
    return train, valid, test, extra


def load_dataset(dataset, dataset_neighbor, target_prop, use_ref=False, use_ring=True):
    """``load_dataset`` (scann/utils/general.py:104-144): the pickled object arrays the reference's preprocessing
    writes.  ``dataset``: records ``{"Atomic": [...], "Properties": {name: value, ...}, "Features": {name: per-atom
    flags}}``; ``dataset_neighbor``: the per-structure neighbour lists.  Returns ``(data_energy, data_neighbor)`` as
    object arrays: rows ``[atomic numbers, target, ring/aromatic flags [atoms, n_features]]`` with ``use_ring``
    (which takes precedence), ``[atomic numbers, target - Ref_energy]`` with ``use_ref``, else
    ``[atomic numbers, target]``."""
    records = np.load(dataset, allow_pickle=True)
    if use_ref:
        print("Using reference energy optimization", "\n")
    if use_ring:
        print("Using ring aromatic information", "\n")
    rows = []
    for rec in records:
        value = float(rec["Properties"][target_prop])
        if use_ring:
            rows.append([rec["Atomic"], value, np.stack([rec["Features"][k] for k in rec["Features"]], -1)])
        elif use_ref:
            rows.append([rec["Atomic"], value - float(rec["Properties"]["Ref_energy"])])
        else:
            rows.append([rec["Atomic"], value])
    data_energy = np.array(rows, dtype="object")
    data_neighbor = np.array(np.load(dataset_neighbor, allow_pickle=True), dtype="object")
    return data_energy, data_neighbor


def prepare_input_from_neighbors(atomic_numbers, neighbors, angle: bool = True) -> Dict[str, np.ndarray]:
    """The padding half of ``prepare_input_pmt`` (scann/utils/general.py:218-246) for a caller who already has the
    neighbour lists of ONE structure -- the README inference flow (README.md:102-120) without pymatgen:
    ``neighbors[a]`` = list of tuples ``n`` as ``compute_voronoi_neighbor`` returns them (``n[1]`` neighbour atom
    index, ``n[2]`` solid angle, ``n[3]`` normalised solid angle, ``n[-1]`` distance).  ``angle=True`` selects the raw
    solid angle (what the g_update models are trained on, general.py:224).  Returns the model's input dict with a
    leading batch axis of 1; padded slots carry index 0 / weight 0 / distance 0 and ``neighbor_mask`` False, exactly
    as ``pad_sequence(..., value=1000)`` + ``!= 1000`` produce them (tests/test_reference_pins.py)."""
    na = len(neighbors)
    if len(atomic_numbers) != na:
        raise ValueError("one neighbour list per atom is required")
    n_max = max((len(lc) for lc in neighbors), default=0)
    nbr = np.full((1, na, n_max), 1000, np.int32)
    w = np.zeros((1, na, n_max), np.float32)
    d = np.zeros((1, na, n_max), np.float32)
    wi = 2 if angle else 3
    for a, lc in enumerate(neighbors):
        k = len(lc)
        if k:
            nbr[0, a, :k] = [n[1] for n in lc]
            w[0, a, :k] = [n[wi] for n in lc]
            d[0, a, :k] = [n[-1] for n in lc]
    mask = nbr != 1000
    nbr[nbr == 1000] = 0
    atomics = np.array([atomic_numbers], "int32")
    return {"atomic": atomics, "atom_mask": np.expand_dims(atomics != 0, -1), "neighbors": nbr, "neighbor_mask": mask,
            "neighbor_weight": w, "neighbor_distance": d}


def prepare_input_pmt(struct, d_t=4.0, w_t=0.4, angle=True, neighbors=None) -> Dict[str, np.ndarray]:
    """``prepare_input_pmt`` (scann/utils/general.py:206-246) with the reference's signature.  The Voronoi neighbour
    search behind it (``compute_voronoi_neighbor``: pymatgen + Qhull) is outside this package, so the caller hands the
    neighbour lists over as ``neighbors`` (what ``compute_voronoi_neighbor(struct, d_thresh=d_t, w_thresh=w_t)`` of the
    reference returns); ``struct`` only has to offer ``atomic_numbers``.  Without ``neighbors`` the call fails loudly."""
    if neighbors is None:
        raise NotImplementedError("prepare_input_pmt: the Voronoi neighbour search (pymatgen / Qhull) is not part of the "
                                  "accelerated package; pass neighbors=compute_voronoi_neighbor(struct, d_thresh=d_t, "
                                  "w_thresh=w_t) or call prepare_input_from_neighbors(atomic_numbers, neighbors)")
    return prepare_input_from_neighbors(list(struct.atomic_numbers), neighbors, angle=angle)


def padded_to_csr(inputs: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Inverse of ``pack_padded`` for batches whose valid atoms / neighbour slots are prefixes (what
    DataIterator.__getitem__ produces): padded dict -> CSR batch for ``Engine.load_batch_csr``."""
    am = inputs["atom_mask"][..., 0].astype(bool)
    nm = inputs["neighbor_mask"].astype(bool)
    B, M, N = nm.shape
    n_at = am.sum(1)
    if not (np.array_equal(am, np.arange(M)[None, :] < n_at[:, None])):
        raise ValueError("valid atoms must be a prefix of every structure")
    n_nb = nm.sum(2)
    if not np.array_equal(nm, np.arange(N)[None, None, :] < n_nb[..., None]):
        raise ValueError("valid neighbour slots must be a prefix of every atom")
    sa = np.zeros(B + 1, np.int32)
    np.cumsum(n_at, out=sa[1:])
    cnt = n_nb[am]
    an = np.zeros(len(cnt) + 1, np.int32)
    np.cumsum(cnt, out=an[1:])
    sel = nm & am[..., None]
    csr = {"struct_atom_off": sa, "atom_nbr_off": an, "z": inputs["atomic"][am].astype(np.int32),
           "nbr_idx": inputs["neighbors"][sel].astype(np.int32), "nbr_w": inputs["neighbor_weight"][sel].astype(np.float32),
           "nbr_d": inputs["neighbor_distance"][sel].astype(np.float32), "M": M, "N": N}
    if "ring_aromatic" in inputs:
        csr["ring"] = inputs["ring_aromatic"][am].astype(np.int32)
    return csr


def synthetic_ragged(n_struct: int, seed: int = 0, max_atoms: int = 29, max_nbr: int = 16, use_ring: bool = False):
    """Random data in the reference's nested-list format (for tests and benchmarks)."""
    rng = np.random.default_rng(seed)
    data_energy, data_neighbor = [], []
    for _ in range(n_struct):
        na = int(rng.integers(2, max_atoms + 1))
        zs = rng.choice([1, 6, 7, 8, 9], size=na).tolist()
        rec = [zs, float(rng.standard_normal())]
        if use_ring:
            rec.append(rng.integers(0, 2, size=(na, 2)).tolist())
        data_energy.append(tuple(rec))
        nb = []
        for _a in range(na):
            k = int(rng.integers(1, max_nbr + 1))
            nb.append([(0, int(rng.integers(0, na)), float(rng.uniform(0.4, 3.0)), float(rng.uniform(0.2, 1.0)),
                        float(rng.uniform(0.9, 4.0))) for _ in range(k)])
        data_neighbor.append(nb)
    de = np.empty(n_struct, object)
    dn = np.empty(n_struct, object)
    for i in range(n_struct):
        de[i], dn[i] = data_energy[i], data_neighbor[i]
    return de, dn
