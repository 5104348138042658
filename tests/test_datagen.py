"""DataIterator on CSR arrays (scann_b200/datagenerator.py) against the loop restatement of the reference's
batch assembly (oracle/datagen_oracle.py; scann/utils/datagenerator.py:69-135): bit-exact."""
import numpy as np
import pytest

from oracle import datagen_oracle as DO
from scann_b200.datagenerator import DataIterator, synthetic_ragged


@pytest.mark.parametrize("g_update,use_ring", [(True, False), (False, True)])
def test_padded_batches_are_bit_identical_to_the_reference_assembly(g_update, use_ring):
    de, dn = synthetic_ragged(37, seed=3, use_ring=use_ring)
    it = DataIterator(de, dn, batch_size=8, use_ring=use_ring, g_update=g_update)
    assert len(it) == 5
    for i in range(len(it)):
        inputs, energy = it[i]
        ref_in, ref_e = DO.get_item(de, dn, list(range(i * 8, min(37, (i + 1) * 8))), g_update=g_update, use_ring=use_ring)
        assert np.array_equal(energy, ref_e)
        assert set(inputs) == set(ref_in)
        for k in ref_in:
            assert inputs[k].dtype == ref_in[k].dtype and inputs[k].shape == ref_in[k].shape, k
            assert np.array_equal(inputs[k], ref_in[k]), k


def test_shuffle_and_last_partial_batch():
    de, dn = synthetic_ragged(10, seed=1)
    np.random.seed(0)
    it = DataIterator(de, dn, batch_size=4, shuffle=True, converter=True)
    order = it.indexes.copy()
    assert sorted(order.tolist()) == list(range(10)) and len(it) == 3
    inputs, energy = it[2]
    assert inputs["atomic"].shape[0] == 2
    ref_in, ref_e = DO.get_item(de, dn, order[8:10].tolist(), g_update=False, converter=1000)
    assert np.array_equal(energy, ref_e) and np.array_equal(inputs["neighbors"], ref_in["neighbors"])


@pytest.mark.parametrize("shape,use_ring", [("qm9", False), ("mp2018", False), ("ptgp", True)])
def test_padded_to_csr_round_trip(shape, use_ring):
    """padded_to_csr (what bench.py uses to hand a padded batch over in ragged form) is the inverse of pack_padded
    for prefix-valid batches, and refuses batches whose valid slots are not prefixes."""
    from scann_b200.datagenerator import pack_padded, padded_to_csr
    from scann_b200.synth import make_batch
    inputs, _ = make_batch(shape, 5, B=6, use_ring=use_ring)
    am = inputs["atom_mask"][..., 0]
    nm = inputs["neighbor_mask"]
    prefix = (np.array_equal(am, np.arange(am.shape[1])[None] < am.sum(1)[:, None]) and
              np.array_equal(nm, np.arange(nm.shape[2])[None, None] < nm.sum(2)[..., None]))
    if not prefix:
        with pytest.raises(ValueError):
            padded_to_csr(inputs)
        return
    csr = padded_to_csr(inputs)
    assert csr["z"].shape[0] == am.sum() and csr["nbr_idx"].shape[0] == (nm & am[..., None]).sum()
    back = pack_padded(csr)
    for k in ("atomic", "atom_mask", "neighbor_mask"):
        assert np.array_equal(back[k], inputs[k]), k
    sel = nm & am[..., None]
    for k in ("neighbors", "neighbor_weight", "neighbor_distance"):
        assert np.array_equal(back[k][sel], inputs[k][sel]), k
    if use_ring:
        assert np.array_equal(back["ring_aromatic"][am], inputs["ring_aromatic"][am])
