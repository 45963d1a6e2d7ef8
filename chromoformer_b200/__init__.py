"""chromoformer_b200 — B200-native (sm_100a) implementation of the Chromoformer hot path.

Public names are the reference's (``chromoformer/__init__.py:1-2``) plus the legacy flat
``Chromoformer`` class of ``chromoformer/net.py:156``.
"""
from .model import (Chromoformer, ChromoformerBase, ChromoformerClassifier,  # noqa: F401
                    ChromoformerRegressor)

__version__ = "0.1.0"
