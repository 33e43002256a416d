"""ecgmm -- B200-native (sm_100a) implementation of the ECG tri-modal fusion classifier hot path.

Import name: ``ecgmm`` (the directory is ``ecg-multimodal-model_b200``; the root-level
``ecgmm.py`` shim maps one onto the other).
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
