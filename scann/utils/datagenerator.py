"""``scann.utils.datagenerator`` (reference: scann/utils/datagenerator.py)."""
from scann_b200.datagenerator import DataIterator  # noqa: F401
