"""B200-native hot path of the Metadata-Augmented U-Net (import name: ``mau_b200``).

Public surface mirrors reference ``src/model.py``: :class:`UrbanPredictor` (same
constructor, ``forward(maps, temp_series, metadata)`` and ``state_dict`` layout).
"""
from .model import UrbanPredictor, UrbanPredictor_unet, UrbanPredictor_unetpp  # noqa: F401
from . import engine  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from . import data  # noqa: F401  (mirror of reference src/dataset.py on the native tile reader)
from . import losses  # noqa: F401  (mirror of reference src/utils/losses.py on the loss kernels)

__all__ = ["UrbanPredictor", "UrbanPredictor_unet", "UrbanPredictor_unetpp", "engine", "FusedAdamW", "data", "losses"]
