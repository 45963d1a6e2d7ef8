"""Loss curves of TrainStep in both precisions on the same 64 genes (diagnostic for tests/test_training_gpu.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from chromoformer_b200 import ChromoformerRegressor, synthetic  # noqa: E402
from chromoformer_b200.trainer import TrainStep  # noqa: E402

KWS = ({"n_layers": 1, "n_heads": 2, "d_model": 128, "d_ff": 128}, {"n_layers": 2, "n_heads": 2, "d_model": 128, "d_ff": 256},
       {"n_layers": 6, "n_heads": 8, "d_model": 256, "d_ff": 256})
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
batch = synthetic.make_batch(64, ragged=True, seed=33)
dev = {k: ({b: t.cuda() for b, t in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in batch.items()}
target = dev["labels_reg"].view(-1, 1)
print("target mean/var", target.mean().item(), target.var().item())
for prec in ("fp32", "bf16"):
    m = ChromoformerRegressor(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=4).cuda().train()
    m.precision = prec
    step = TrainStep(m, lr=lr, regression=True, use_graph=False)
    curve = [float(step(dev, target).item()) for _ in range(steps)]
    print(prec, " ".join(f"{x:.4f}" for x in curve))
