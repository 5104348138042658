"""``from scann.models import SCANN`` -- same import path as the reference (scann/models/__init__.py)."""
from scann_b200.model import SCANN, create_model, ScannKerasModel, CosineDecay  # noqa: F401

__all__ = ["SCANN", "create_model"]
