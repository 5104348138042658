"""Reference layer names (scann/layers/__init__.py:7-17) backed by the sm_100a kernels."""
from scann_b200.layers import (GaussianExpansion, GlobalAttention, LocalAttention, ResidualNorm,  # noqa: F401
                               gather_shape, mrelu, r2_square, root_mean_squared_error)

_CUSTOM_OBJECTS = globals()

__all__ = ["GlobalAttention", "LocalAttention", "ResidualNorm", "GaussianExpansion", "root_mean_squared_error",
           "r2_square", "gather_shape", "mrelu"]
