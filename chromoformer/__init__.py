"""Drop-in import surface of the reference package (``chromoformer/__init__.py:1-2``), backed by
``chromoformer_b200``: ``from chromoformer import ChromoformerClassifier, ChromoformerRegressor,
ChromoformerDataset`` keeps working for demo/run_demo.py and friends."""
from chromoformer_b200.model import ChromoformerClassifier, ChromoformerRegressor  # noqa: F401
from chromoformer_b200.data import ChromoformerDataset  # noqa: F401
