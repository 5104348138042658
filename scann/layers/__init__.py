"""Reference layer names (scann/layers/__init__.py:1-17) backed by the sm_100a kernels."""
from scann_b200.callbacks import SGDRC  # noqa: F401  (the reference exports its lr callback from the layers package)
from scann_b200.layers import (GaussianExpansion, GlobalAttention, LocalAttention, ResidualNorm,  # noqa: F401
                               gather_shape, mrelu, r2_square, root_mean_squared_error)

_CUSTOM_OBJECTS = globals()

__all__ = ["GlobalAttention", "LocalAttention", "ResidualNorm", "GaussianExpansion", "SGDRC", "root_mean_squared_error",
           "r2_square", "gather_shape", "mrelu"]
