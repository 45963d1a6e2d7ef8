"""Tiny driver for ncu: two forward passes over one chunk of synthetic genes (first = warm-up)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from chromoformer_b200 import ChromoformerClassifier, ChromoformerRegressor, synthetic  # noqa: E402
from chromoformer_b200.engine import InferenceEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "infer"
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
ragged = len(sys.argv) > 4 and sys.argv[4] == "ragged"
if mode == "infer":
    model = ChromoformerClassifier(seed=123).cuda().eval()
    model.precision = precision
    eng = InferenceEngine(model, chunk=n)
    batch = eng.to_device(synthetic.make_batch(n, ragged=ragged, seed=0))
    for _ in range(2):
        eng.predict_device(batch)
else:
    from chromoformer_b200.trainer import TrainStep
    model = ChromoformerRegressor(seed=123).cuda().train()
    model.precision = precision
    tb = synthetic.make_batch(n, ragged=False, seed=0)
    dev = {k: ({b: t.cuda() for b, t in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in tb.items()}
    step = TrainStep(model, regression=True)
    for _ in range(2):
        step(dev, dev["labels_reg"].view(-1, 1))
torch.cuda.synchronize()
print("done")
