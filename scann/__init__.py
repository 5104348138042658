"""Import-compatible alias of the reference package name: ``from scann.models import SCANN``.
Everything is implemented in :mod:`scann_b200` (sm_100a kernels behind a C ABI)."""
