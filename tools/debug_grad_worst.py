import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from _util import KWS
from chromoformer_b200 import ChromoformerRegressor, synthetic
from oracle import chromoformer_oracle as oracle
model = ChromoformerRegressor(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=31)
sd = {k: v.detach().cpu().clone() for k, v in model.named_parameters()}
batch = synthetic.make_batch(70, i_max=8, w_max=20000, n_feats=7, ragged=True, full_masks=True, seed=41)
target = batch["labels_reg"].view(-1, 1)
loss_o, logits_o, grads_o = oracle.forward_backward(sd, synthetic.forward_args(batch), target, True)
model.cuda().train()
out = model(*synthetic.forward_args(batch, "cuda"))
torch.nn.functional.mse_loss(out, target.cuda()).backward()
rows = []
for name, p in model.named_parameters():
    g = grads_o[name]
    if g is None: continue
    rows.append(((p.grad.cpu() - g).abs().max().item() / max(g.abs().max().item(), 1e-7), name, g.abs().max().item()))
for r in sorted(rows, reverse=True)[:8]: print(r)
