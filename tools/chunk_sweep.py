"""genes/s of the device-resident 18,955-gene sweep (BF16) as a function of the chunk size of InferenceEngine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from chromoformer_b200 import ChromoformerClassifier, synthetic  # noqa: E402
from chromoformer_b200.engine import InferenceEngine  # noqa: E402

N = 18955
model = ChromoformerClassifier(seed=123).cuda().eval()
model.precision = "bf16"
host = synthetic.make_batch(N, ragged=False, seed=0)
chunks = [int(c) for c in sys.argv[1:]] or [4096, 4144, 6216, 9478, 18955]
res = None
for c in chunks:
    eng = InferenceEngine(model, chunk=c)
    if res is None:
        res = eng.to_device(host)
    out = torch.empty(N, 2, device="cuda")
    for _ in range(3):
        eng.predict_device(res, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        eng.predict_device(res, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"chunk {c}: {ms:.3f} ms/sweep = {N / ms * 1e3 / 1e6:.3f} M genes/s")
    model._ws_cache.clear()
