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
