"""``chromoformer.util.seed_everything`` (reference util.py:6-12)."""
import os
import random

import numpy as np
import torch


def seed_everything(seed=42):
    os.environ["PYTHONHASHSEED"] = str(seed)
    for fn in (random.seed, np.random.seed, torch.manual_seed):
        fn(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = True
