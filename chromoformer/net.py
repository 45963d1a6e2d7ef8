"""``chromoformer.net`` names (reference net.py) served by the sm_100a implementation."""
from chromoformer_b200.model import (Chromoformer, ChromoformerBase, ChromoformerClassifier,  # noqa: F401
                                     ChromoformerRegressor, EmbeddingTransformer,
                                     PairwiseInteractionTransformer, RegulationTransformer)
