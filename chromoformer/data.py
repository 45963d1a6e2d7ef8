"""``chromoformer.data`` names (reference data.py) served by chromoformer_b200.data."""
from chromoformer_b200.data import ChromoformerDataset, GeneBatcher, bin_regions_device  # noqa: F401
