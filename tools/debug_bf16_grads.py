import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chromoformer_b200 import ChromoformerRegressor, synthetic
from oracle import chromoformer_oracle as oracle
n = 64
model = ChromoformerRegressor(seed=11)
sd = {k: v.detach().clone() for k, v in model.named_parameters()}
batch = synthetic.make_batch(n, ragged=True, full_masks=False, seed=21)
target = batch["labels_reg"].view(-1, 1)
loss_o, logits_o, grads_o = oracle.forward_backward(sd, synthetic.forward_args(synthetic.expand_full_masks(batch)), target, True)
model.cuda().train()
model.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
out = model(*synthetic.forward_args(batch, "cuda"))
loss = torch.nn.functional.mse_loss(out, target.cuda())
loss.backward()
rows = []
for name, p in model.named_parameters():
    g = grads_o[name]
    if g is None: continue
    sc = max(g.abs().max().item(), 1e-9)
    rows.append(((p.grad.cpu() - g).abs().max().item() / sc, name, sc))
rows.sort(reverse=True)
print("loss", loss.item(), loss_o.item())
for r in rows[:14]: print("%.3e  %-70s scale %.2e" % r)
print("median", rows[len(rows)//2][0])
