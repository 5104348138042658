"""``scann.utils`` import path of the reference (scann/utils/__init__.py) on the accelerated package: the data
iterator, the padding helpers, the data-set loaders.  Voronoi neighbour search and the file loaders (pymatgen,
openbabel) are outside the accelerated path (DESIGN.md section 1, row f-N4)."""
from scann_b200.datagenerator import (DataIterator, load_atomic_features, load_dataset, pad_nested_sequences,  # noqa: F401
                                      pad_sequence, prepare_input_from_neighbors, prepare_input_pmt, set_atomic_features,
                                      split_data)

__all__ = ["DataIterator", "pad_sequence", "pad_nested_sequences", "split_data", "load_dataset",
           "prepare_input_pmt", "prepare_input_from_neighbors", "set_atomic_features", "load_atomic_features"]
