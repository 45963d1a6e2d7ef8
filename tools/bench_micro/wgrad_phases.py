"""Clocks of CTA 0 of the grouped weight-gradient launch of one training step (chromo_debug_trace)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from chromoformer_b200 import ChromoformerRegressor, _lib, synthetic  # noqa: E402
from chromoformer_b200.trainer import TrainStep  # noqa: E402

lib = _lib.load()
reg = ChromoformerRegressor(seed=123).cuda().train()
reg.precision = "bf16"
tb = synthetic.make_batch(64, ragged=False, seed=100)
dev = {k: ({b: t.cuda() for b, t in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in tb.items()}
target = dev["labels_reg"].view(-1, 1)
step = TrainStep(reg, lr=3e-5, regression=True, use_graph=False)
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
for i in range(3):
    lib.chromo_debug_trace(buf.data_ptr() if i == 2 else None)
    step(dev, target)
    torch.cuda.synchronize()
lib.chromo_debug_trace(None)
t = buf[3072:3136].cpu().tolist()
n = t[63]
mask = 0xfffffffff
t0 = t[0] & mask
print("events", n)
i = 1
while i + 2 < n:
    tok, nc = t[i] >> 48, (t[i] >> 36) & 0xfff
    a, b, c = t[i] & mask, t[i + 1] & mask, t[i + 2] & mask
    print(f"item tokens={tok} n_cols={nc}: start {a - t0:8d}  contract {b - a:7d}  epilogue {c - b:6d}")
    i += 3
print("end", (t[n - 1] & mask) - t0)
