"""``scann.utils.general`` (reference: scann/utils/general.py) -- the numpy half."""
from scann_b200.datagenerator import (load_dataset, pad_nested_sequences, pad_sequence,  # noqa: F401
                                      prepare_input_from_neighbors, prepare_input_pmt, split_data)
