"""splice_b200 — B200-native (sm_100a) hot path of omerbt/Splice behind the reference's Python entry points."""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_MIRRORED = ("models", "models.model", "models.extractor", "models.networks", "models.unet", "models.unet.skip",
             "models.unet.common", "util", "util.losses", "util.util", "data", "data.Dataset", "data.transforms")


def install_as_reference_modules() -> None:
    """Register the mirrors under the reference's own module names (`models.model`, `util.losses`, ...), so that the
    reference's unmodified `train.py` / notebook import the sm_100a implementations."""
    for name in _MIRRORED:
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")
