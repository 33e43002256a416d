"""ecgmm -- B200-native (sm_100a) implementation of the ECG tri-modal fusion classifier hot path.

Import name: ``ecgmm`` (the directory is ``ecg-multimodal-model_b200``; the root-level
``ecgmm.py`` shim maps one onto the other).

    from ecgmm import ECGMultimodalModel, nn as enn, optim as eoptim
    model = ECGMultimodalModel(Config)            # same constructor / forward / state_dict as the reference
    criterion = enn.CrossEntropyLoss()            # train.py:31
    optimizer = eoptim.Adam(model.parameters(), lr=Config.learning_rate)   # train.py:43
"""
from . import lib  # noqa: F401
from . import ops  # noqa: F401
from . import data, explain, graph, model, nn, optim, preprocess, serve  # noqa: F401
from .model import ECGMultimodalModel, FusionClassifierWrapper, MultimodalModel, ResNet1D_SE, ResNet18  # noqa: F401
from .nn import CrossEntropyLoss, FocalLoss  # noqa: F401

__all__ = ["lib", "ops", "data", "model", "nn", "optim", "preprocess", "explain", "graph", "serve", "ECGMultimodalModel", "MultimodalModel", "ResNet1D_SE", "ResNet18",
           "FusionClassifierWrapper", "CrossEntropyLoss", "FocalLoss"]
